# Round 2, GPU call 20: final dispatch (forced operators back on the direct kernel by default), NVML clock sampler in bench.py
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_native_step_more_gpu.py tests/test_native_step_gpu.py -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c20_pytest.log; tail -4 gpurun_out/r2c20_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-secondary --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2c20_bench_short.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-secondary --no-cpu-baseline --force 1e-5 2>&1 | tail -1 | tee gpurun_out/r2c20_bench_forced.json
