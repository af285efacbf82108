# Round 2, GPU call 3: exact-arithmetic build — full GPU suite (bit identity vs the C oracle), perf impact, new bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rfEs -s 2>&1 | grep -v "^registered bc\|^$" > gpurun_out/r2c3_pytest.log; tail -15 gpurun_out/r2c3_pytest.log
: > gpurun_out/r2c3_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2c3_matrix.log; }
run
run --config periodic
run --policy FP32FP16
run --policy FP32FP16 --config periodic
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --collision KBC --config sphere
run --policy FP64FP32
run --policy FP64FP64 --n 384
run --n 128 --steps 1000
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c3_bench_default.json 2> gpurun_out/r2c3_bench_default.err; tail -c 3000 gpurun_out/r2c3_bench_default.json
