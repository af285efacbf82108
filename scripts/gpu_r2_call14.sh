# Round 2, GPU call 14: scalar tile kernel as the D3Q19 FP32FP32 default: whole GPU suite, default bench line, launch list
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c14_pytest.log; tail -8 gpurun_out/r2c14_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c14_bench_default.json 2> gpurun_out/r2c14_bench_default.err; tail -c 1500 gpurun_out/r2c14_bench_default.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c14_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c14_launch_bench.log 2>&1
python __graft_entry__.py smoke > gpurun_out/r2c14_smoke.log 2>&1 || python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c14_smoke.log 2>&1; tail -3 gpurun_out/r2c14_smoke.log
