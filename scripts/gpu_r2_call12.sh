# Round 2, GPU call 12: 1024-cell tiles as the default where they fit; whole suite; final bench line and ncu capture of the shipped tile kernel
set -x
mkdir -p gpurun_out
: > gpurun_out/r2c12_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c12_matrix.log; }
run --policy FP32FP16
run --policy FP32FP16 --config periodic
run --policy FP32FP16 --config sphere
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27 --policy FP32FP16 --config periodic
run --policy FP32FP16 --cells-per-thread 404
run --policy FP32FP16 --n 256 --steps 100
run --policy FP32FP16 --n 128 --steps 500
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c12_pytest.log; tail -8 gpurun_out/r2c12_pytest.log
ncu --set full --clock-control none --import-source on -k regex:step_tile -s 4 -c 1 -o gpurun_out/r2c12_tile python bench.py --policy FP32FP16 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c12_ncu.log 2>&1
ncu -i gpurun_out/r2c12_tile.ncu-rep --page raw --csv > gpurun_out/r2c12_tile_raw.csv 2>/dev/null; rm -f gpurun_out/r2c12_tile.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c12_bench_default.json 2> gpurun_out/r2c12_bench_default.err; tail -c 1200 gpurun_out/r2c12_bench_default.json
