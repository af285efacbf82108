# Round 2, GPU call 21: ncu evidence of the final headline kernel (launch list of the bench command + one --set full capture with source)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c21_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c21_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_tile1 -s 4 -c 1 -o gpurun_out/r2c21_t1 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c21_ncu.log 2>&1
ncu -i gpurun_out/r2c21_t1.ncu-rep --page raw --csv > gpurun_out/r2c21_t1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c21_t1.ncu-rep --page source --csv > gpurun_out/r2c21_t1_source.csv 2>/dev/null
ls -la gpurun_out/r2c21_t1.ncu-rep; rm -f gpurun_out/r2c21_t1.ncu-rep
