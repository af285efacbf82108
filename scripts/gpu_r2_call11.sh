# Round 2, GPU call 11 (tuning experiment): tile kernel with 1024-cell tiles — 16 consumer warps and one producer per CTA, one CTA per SM, 2 KB bulk copies
set -x
mkdir -p gpurun_out
: > gpurun_out/r2c11_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$(basename ${XLB_B200_LIB:-default}) $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c11_matrix.log; }
run --policy FP32FP16
run --policy FP32FP16 --config periodic
export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_t1024.so
timeout 300 python -m pytest tests/test_native_step_more_gpu.py -m gpu -q -p no:cacheprovider -x -k "tile_kernel_is_bit" 2>&1 | tail -3 | tee gpurun_out/r2c11_tests.log
run --policy FP32FP16
run --policy FP32FP16 --config periodic
run --policy FP32FP16 --config sphere
