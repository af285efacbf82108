set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
for v in 1 2 4; do
  timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
done
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --config periodic 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
for v in 2 4 8; do
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --policy FP32FP16 --cells-per-thread $v 2>&1 | tail -1 | tee -a gpurun_out/bench_v.log
done
