# Round-1 record run: GPU test-suite, smoke, the default bench line, ncu launch list + full capture of the step kernel.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
XLB_FULL_C1=1 timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/prof_r1_d3q19_f32 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
