# Multi-GPU record run (gpurun --gpus 8): bit-identity check at 8 and 4 slabs, weak scaling of the default bench, C4 and fp16.
mkdir -p gpurun_out; : > gpurun_out/scaling.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
$TR --nproc-per-node 8 --master-port 29601 scripts/mgpu_check.py 2>&1 | grep -E "mgpu|MGPU" | tee -a gpurun_out/scaling.log
$TR --nproc-per-node 4 --master-port 29602 scripts/mgpu_check.py 2>&1 | grep -E "MGPU" | tee -a gpurun_out/scaling.log
port=29610
run() { n=$1; shift; port=$((port+1)); out=$(timeout 600 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 40 --warmup 5 --no-cpu-baseline "$@" 2>&1 | grep '^{' | tail -1); echo "$out" >> gpurun_out/scaling_lines.jsonl; echo "N=$n $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', (d.get('e2e') or {}).get('value'))" 2>/dev/null || echo FAILED)" | tee -a gpurun_out/scaling.log; }
: > gpurun_out/scaling_lines.jsonl
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | grep '^{' | tail -1 | tee -a gpurun_out/scaling_lines.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 =>', d['value'], d['ms_per_step'], d['roofline']['frac'])" | tee -a gpurun_out/scaling.log
run 2 --no-e2e
run 4 --no-e2e
run 8 --no-e2e
run 8 --no-e2e --policy FP32FP16
run 8 --no-e2e --lattice D3Q27 --collision KBC --config tunnel
run 8 --no-e2e --lattice D3Q27 --collision KBC
run 8 --e2e-steps 20
