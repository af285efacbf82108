# Round 2, GPU call 6: tile kernel v3 (EquilibriumBC constants in shared memory folded into the select, warp rotation), exact KBC, reference scripts
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_native_step_more_gpu.py -m gpu -q -p no:cacheprovider -x -k "tile_kernel" 2>&1 | tail -5 | tee gpurun_out/r2c6_tile_tests.log
: > gpurun_out/r2c6_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c6_matrix.log; }
run --policy FP32FP16 --cells-per-thread 402
run --policy FP32FP16 --config periodic --cells-per-thread 402
run --lattice D3Q27 --policy FP32FP16 --cells-per-thread 402
run --lattice D3Q27 --policy FP32FP16 --config periodic --cells-per-thread 402
run --policy FP32FP16 --config sphere --cells-per-thread 402
run --policy FP32FP16 --n 256 --steps 100 --cells-per-thread 402
run --policy FP32FP16 --n 128 --steps 500 --cells-per-thread 402
run --lattice D3Q27 --collision KBC --cells-per-thread 300
timeout 900 python -m pytest tests/test_reference_scripts_gpu.py tests/test_native_long_runs_gpu.py tests/test_native_step_more_gpu.py tests/test_native_slab_gpu.py tests/test_examples_gpu.py -m gpu -q -p no:cacheprovider -rfEs -s 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c6_pytest.log; tail -12 gpurun_out/r2c6_pytest.log
ncu --set full --clock-control none --import-source on -k regex:step_tile -s 4 -c 1 -o gpurun_out/r2c6_tile python bench.py --policy FP32FP16 --cells-per-thread 402 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r2c6_ncu.log 2>&1
ncu -i gpurun_out/r2c6_tile.ncu-rep --page raw --csv > gpurun_out/r2c6_tile_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c6_tile.ncu-rep --page source --csv > gpurun_out/r2c6_tile_source.csv 2>/dev/null; rm -f gpurun_out/r2c6_tile.ncu-rep
