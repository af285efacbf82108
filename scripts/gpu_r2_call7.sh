# Round 2, GPU call 7: everything as it will be judged — whole GPU suite, smoke, the default bench line with its secondary workloads, the reference arm
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2c7_smoke.log
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c7_pytest.log; tail -12 gpurun_out/r2c7_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c7_bench_default.json 2> gpurun_out/r2c7_bench_default.err; tail -c 4000 gpurun_out/r2c7_bench_default.json
: > gpurun_out/r2c7_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c7_matrix.log; }
run
run --config periodic
run --lattice D3Q27
run --lattice D3Q27 --collision KBC
run --policy FP64FP32
run --n 128 --steps 1000
run --n 256 --steps 200
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c7_bench_reference.json 2> gpurun_out/r2c7_bench_reference.err; tail -c 1500 gpurun_out/r2c7_bench_reference.json
