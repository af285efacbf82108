# Round 2, GPU call 10: the streamed host job (stepper.run_streamed) — bit identity, then the e2e figure of the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_native_step_more_gpu.py -m gpu -q -p no:cacheprovider -x -k "streamed or graph" 2>&1 | tail -15 | tee gpurun_out/r2c10_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/r2c10_bench_default.json 2> gpurun_out/r2c10_bench_default.err; tail -c 2500 gpurun_out/r2c10_bench_default.json; tail -5 gpurun_out/r2c10_bench_default.err
timeout 900 python bench.py --steps 100 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2c10_bench_100.json 2> gpurun_out/r2c10_bench_100.err; tail -c 1200 gpurun_out/r2c10_bench_100.json
timeout 900 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --policy FP32FP16 > gpurun_out/r2c10_bench_fp16.json 2> gpurun_out/r2c10_bench_fp16.err; tail -c 1200 gpurun_out/r2c10_bench_fp16.json
