# full GPU test-suite + benchmark matrix over lattices / policies with the product library
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest.log
: > gpurun_out/matrix.log
run() { out=$(timeout 400 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/matrix.log; }
run
run --config periodic
for v in 1 2 4; do run --policy FP32FP16 --cells-per-thread $v; done
run --policy FP32FP16 --config periodic
run --policy FP64FP32
run --policy FP64FP64 --n 384
run --lattice D3Q27
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27 --collision KBC --policy FP32FP16
run --n 128
run --n 256
