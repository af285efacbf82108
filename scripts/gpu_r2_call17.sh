# Round 2, GPU call 17 (8 GPUs, trimmed): scaling points with the scalar tile kernel on the slab interiors, slab bit-identity at 8
set -x
mkdir -p gpurun_out
: > gpurun_out/r2c17_scaling.txt
: > gpurun_out/r2c17_scaling_lines.jsonl
run() { n=$1; shift; out=$(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | grep '^{' | tail -1); echo "$out" >> gpurun_out/r2c17_scaling_lines.jsonl; echo "N=$n $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['scaling'])" 2>/dev/null || echo FAILED)" | tee -a gpurun_out/r2c17_scaling.txt; }
run 8
run 8 --scaling strong
run 4 --scaling strong
run 4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/mgpu_check.py 2>&1 | grep -v "^registered" | tail -13 | tee gpurun_out/r2c17_mgpu_check.txt
