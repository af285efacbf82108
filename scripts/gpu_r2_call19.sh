# Round 2, GPU call 19: extended collision operators (D3Q19) on the scalar tile kernel; whole suite with the final library; default bench line
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c19_pytest.log; tail -8 gpurun_out/r2c19_pytest.log
: > gpurun_out/r2c19_matrix.log
run() { out=$(timeout 300 python bench.py --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c19_matrix.log; }
run --steps 20 --collision SmagorinskyLESBGK
run --steps 20 --collision SmagorinskyLESBGK --cells-per-thread 1
run --steps 20 --force 1e-5
run --steps 20 --force 1e-5 --cells-per-thread 1
run --steps 20 --collision SmagorinskyLESBGK --force 1e-5
run --steps 20 --collision SmagorinskyLESBGK --force 1e-5 --cells-per-thread 1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c19_bench_default.json 2> gpurun_out/r2c19_bench_default.err; tail -c 400 gpurun_out/r2c19_bench_default.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c19_smoke.log 2>&1; tail -3 gpurun_out/r2c19_smoke.log
