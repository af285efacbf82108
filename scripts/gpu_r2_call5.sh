# Round 2, GPU call 5: tile kernel v2 (direct stores, no CTA barrier; 2 vs 3 CTAs per SM), full suite incl. 2-D slabs and the reference's own scripts
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_native_step_more_gpu.py -m gpu -q -p no:cacheprovider -x -k "tile_kernel" 2>&1 | tail -5 | tee gpurun_out/r2c5_tile_tests.log
: > gpurun_out/r2c5_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c5_matrix.log; }
run --policy FP32FP16 --cells-per-thread 202
run --policy FP32FP16 --cells-per-thread 402
run --policy FP32FP16 --cells-per-thread 403
run --policy FP32FP16 --config periodic --cells-per-thread 402
run --policy FP32FP16 --config periodic --cells-per-thread 403
run --lattice D3Q27 --policy FP32FP16 --cells-per-thread 402
run --lattice D3Q27 --policy FP32FP16 --config periodic --cells-per-thread 402
run --policy FP32FP16 --config sphere --cells-per-thread 402
run --policy FP32FP16 --config sphere --cells-per-thread 403
run --policy FP32FP16 --n 128 --steps 500 --cells-per-thread 402
run --policy FP32FP16 --n 256 --steps 100 --cells-per-thread 403
run --policy FP64FP32
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c5_pytest.log; tail -25 gpurun_out/r2c5_pytest.log
for v in 402 403; do
ncu --set full --clock-control none --import-source on -k regex:step_tile -s 4 -c 1 -o gpurun_out/r2c5_tile$v python bench.py --policy FP32FP16 --cells-per-thread $v --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c5_ncu$v.log 2>&1
ncu -i gpurun_out/r2c5_tile$v.ncu-rep --page raw --csv > gpurun_out/r2c5_tile${v}_raw.csv 2>/dev/null
rm -f gpurun_out/r2c5_tile$v.ncu-rep
done
