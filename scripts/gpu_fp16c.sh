mkdir -p gpurun_out; : > gpurun_out/fp16c.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest.log
run() { out=$(timeout 400 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" 2>/dev/null || echo "FAILED: $out" | cut -c1-300)" | tee -a gpurun_out/fp16c.log; }
run --policy FP32FP16
run --policy FP32FP16 --config periodic
run --policy FP32FP16 --lattice D3Q27
run --policy FP32FP16 --lattice D3Q27 --config periodic
run --policy FP32FP16 --n 256
run --policy FP32FP16 --n 128
run
