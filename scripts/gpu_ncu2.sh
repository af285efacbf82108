mkdir -p gpurun_out
cap() { name=$1; shift; timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 3 -c 1 -o /tmp/$name python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /tmp/$name.log 2>&1; ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; }
cap prof_r1_d3q19_f16_v1 --policy FP32FP16 --cells-per-thread 1
cap prof_r1_d3q19_f16_v2 --policy FP32FP16 --cells-per-thread 2
cap prof_r1_d3q19_f16_v102 --policy FP32FP16 --cells-per-thread 102
cap prof_r1_d3q27_kbc_f32 --lattice D3Q27 --collision KBC
cap prof_r1_d3q19_f64f32 --policy FP64FP32
ls -la gpurun_out/*.raw.csv
