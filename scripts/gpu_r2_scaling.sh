# Round 2, multi-GPU call (one 8-GPU box):  gpurun --gpus 8 --timeout 1500 -- 'bash scripts/gpu_r2_scaling.sh'
# strong scaling (fixed global grid split into x-slabs) at N = 2, 4, 8 for the 512^3 D3Q19 BGK cavity and the C3 sphere (D3Q27 KBC
# 1024x512x512), C4 (bluff body, 8 GPUs), the weak-scaling point at 8 for reference, and the slab bit-identity check
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_scaling_lines.jsonl
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
run() { n=$1; shift; out=$(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | grep '^{' | tail -1); echo "$out" >> gpurun_out/r2_scaling_lines.jsonl; echo "N=$n $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['scaling'])" 2>/dev/null || echo FAILED)" | tee -a gpurun_out/r2_scaling.txt; }
: > gpurun_out/r2_scaling.txt
for n in 2 4 8; do run $n --scaling strong; done
for n in 2 4 8; do run $n --config sphere --lattice D3Q27 --collision KBC; done
run 8 --config tunnel --lattice D3Q27 --collision KBC
run 8
run 8 --policy FP32FP16
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 scripts/mgpu_check.py 2>&1 | tail -12 | tee gpurun_out/r2_mgpu_check.txt
