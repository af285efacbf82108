# Round 2, GPU call 13: the scalar tile kernel (cells_per_thread 501 / 502): parity tests, then the matrix against the direct kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_native_step_more_gpu.py -m gpu -q -p no:cacheprovider -rfEs -k "scalar_tile" 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c13_pytest.log; tail -12 gpurun_out/r2c13_pytest.log
: > gpurun_out/r2c13_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c13_matrix.log; }
for v in 0 501 502; do
run --cells-per-thread $v
run --cells-per-thread $v --config periodic
run --cells-per-thread $v --n 128 --steps 500
run --cells-per-thread $v --n 256 --steps 100
run --cells-per-thread $v --lattice D3Q27
done
for v in 0 501; do
run --cells-per-thread $v --policy FP64FP32
run --cells-per-thread $v --policy FP64FP32 --config sphere
run --cells-per-thread $v --policy FP64FP32 --n 128 --steps 300
done
ncu --set full --clock-control none --import-source on -k regex:step_tile1 -s 4 -c 1 -o gpurun_out/r2c13_t1 python bench.py --cells-per-thread 501 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c13_ncu.log 2>&1
ncu -i gpurun_out/r2c13_t1.ncu-rep --page raw --csv > gpurun_out/r2c13_t1_raw.csv 2>/dev/null; rm -f gpurun_out/r2c13_t1.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:step_tile1 -s 4 -c 1 -o gpurun_out/r2c13_t1d python bench.py --cells-per-thread 501 --policy FP64FP32 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c13_ncud.log 2>&1
ncu -i gpurun_out/r2c13_t1d.ncu-rep --page raw --csv > gpurun_out/r2c13_t1d_raw.csv 2>/dev/null; rm -f gpurun_out/r2c13_t1d.ncu-rep
