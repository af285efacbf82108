# Round 2, GPU call 15: scalar tile kernel with its own straight-line store path; step call without tensor-subclass dispatch overhead (small grids)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_native_step_more_gpu.py tests/test_native_long_runs_gpu.py -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c15_pytest.log; tail -5 gpurun_out/r2c15_pytest.log
: > gpurun_out/r2c15_matrix.log
run() { out=$(timeout 300 python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c15_matrix.log; }
run --steps 30
run --steps 30 --config periodic
run --steps 100 --n 256
run --steps 500 --n 128
run --steps 500 --n 128 --cells-per-thread 1
run --steps 1000 --n 64
run --steps 500 --n 128 --policy FP32FP16
run --steps 500 --n 128 --policy FP64FP32
run --steps 500 --n 128 --lattice D3Q27 --collision KBC
run --steps 30 --policy FP64FP32 --cells-per-thread 501
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c15_launches_128.csv python bench.py --n 128 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c15_launch_bench.log 2>&1
