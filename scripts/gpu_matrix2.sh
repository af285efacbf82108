set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest.log
: > gpurun_out/matrix2.log
run() { out=$(timeout 600 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/matrix2.log; }
run
for v in 1 2 102; do run --policy FP32FP16 --cells-per-thread $v; done
run --policy FP32FP16 --config periodic
for v in 1 102; do run --lattice D3Q27 --collision KBC --cells-per-thread $v; done
for v in 1 102; do run --lattice D3Q27 --collision KBC --policy FP32FP16 --cells-per-thread $v; done
for v in 1 102; do run --lattice D3Q27 --policy FP32FP16 --cells-per-thread $v; done
run --cells-per-thread 102
run --lattice D3Q27 --collision KBC --config sphere
run --lattice D3Q27 --collision KBC --config tunnel --steps 20
