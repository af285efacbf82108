# Round 2, GPU call 22: fp16 tile kernel on the scalar copy plan: step + slab + long-run suites, fp16 points
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_native_step_more_gpu.py tests/test_native_slab_gpu.py tests/test_native_long_runs_gpu.py tests/test_native_fullsize_gpu.py -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c22_pytest.log; tail -4 gpurun_out/r2c22_pytest.log
: > gpurun_out/r2c22_matrix.log
run() { out=$(timeout 300 python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['clocks'].get('samples'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c22_matrix.log; }
run --steps 30 --policy FP32FP16
run --steps 30 --policy FP32FP16 --config periodic
run --steps 30 --policy FP32FP16 --lattice D3Q27
run --steps 30 --policy FP32FP16 --cells-per-thread 404
