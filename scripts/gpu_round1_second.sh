set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
: > gpurun_out/bench_v.log
for v in 1 2 4; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --config periodic --cells-per-thread 1 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
for v in 1 2 4 8; do
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --policy FP32FP16 --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
done
for v in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --lattice D3Q27 --collision KBC --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --lattice D3Q27 --collision BGK --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
done
# ncu: launch list of the default bench command, then one full capture of the step kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 2 -o gpurun_out/prof_d3q19_f32_v1 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --cells-per-thread 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/prof_d3q19_f16_v2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --policy FP32FP16 --cells-per-thread 2 > gpurun_out/ncu_full16.log 2>&1
ls -la gpurun_out
