# Round-1 final record run on one B200
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
XLB_FULL_C1=1 timeout 2400 python -m pytest tests -m gpu -x -q --durations=3 2>&1 | tail -10 | tee gpurun_out/pytest.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 900 python bench.py --policy FP32FP16 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp16.json
timeout 900 python bench.py --lattice D3Q27 --collision KBC --config sphere --steps 40 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_c3_sphere.json
timeout 600 python bench.py --n 128 --steps 1000 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_128.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > /tmp/ncu_launch.log 2>&1
cap() { name=$1; shift; timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o /tmp/$name python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /tmp/$name.log 2>&1; ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; }
cap prof_r1_d3q19_f32
cap prof_r1_d3q19_f16_h2 --policy FP32FP16
ls -la gpurun_out
