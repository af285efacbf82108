# Round 2, GPU call 2: the whole GPU suite with the promoted / new tests, then the fp16 and KBC A/B timings
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs -s 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c2_pytest.log; tail -40 gpurun_out/r2c2_pytest.log
: > gpurun_out/r2c2_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$XLB_B200_LIB $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c2_matrix.log; }
run
run --policy FP32FP16
run --policy FP32FP16 --config periodic
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27 --policy FP32FP16 --config periodic
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --collision KBC --cells-per-thread 300
run --policy FP32FP16 --config sphere
for mb in 9 10; do
  export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_h2mb$mb.so
  run --policy FP32FP16
  run --policy FP32FP16 --config periodic
done
unset XLB_B200_LIB
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r2c2_h2 python bench.py --policy FP32FP16 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2c2_ncu_h2.log 2>&1
ncu -i gpurun_out/r2c2_h2.ncu-rep --page raw --csv > gpurun_out/r2c2_h2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c2_h2.ncu-rep --page source --csv > gpurun_out/r2c2_h2_source.csv 2>/dev/null; rm -f gpurun_out/r2c2_h2.ncu-rep
