# Round 2, GPU call 8 (final evidence): ncu captures of the three dominant kernels as shipped, the launch list of the default bench,
# the KBC figures with the lean unit back on -fmad=true, and the whole GPU suite once more
set -x
mkdir -p gpurun_out
: > gpurun_out/r2c8_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c8_matrix.log; }
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --collision KBC --config sphere
run --lattice D3Q27 --collision KBC --policy FP32FP16
run --policy FP64FP32
run --policy FP32FP16
run --policy FP32FP16 --config sphere
run --lattice D3Q27 --collision KBC --config periodic --force 1e-6
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c8_pytest.log; tail -8 gpurun_out/r2c8_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_default.csv python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c8_launches.log 2>&1
cap() { name=$1; shift; ncu --set full --clock-control none --import-source on -k regex:step_ -s 4 -c 1 -o gpurun_out/$name python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary "$@" > gpurun_out/$name.log 2>&1; ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null; rm -f gpurun_out/$name.ncu-rep; }
cap r2_ncu_full_step_kernel_d3q19_f32
cap r2_ncu_tile_final --policy FP32FP16
cap r2_ncu_full_step_kernel_d3q27_kbc_lean_512 --lattice D3Q27 --collision KBC
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c8_bench_default.json 2> gpurun_out/r2c8_bench_default.err; tail -c 1500 gpurun_out/r2c8_bench_default.json
