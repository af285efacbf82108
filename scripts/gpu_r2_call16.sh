# Round 2, GPU call 16 (2 GPUs): slab bit-identity with the tile kernels on the interior launches, strong / weak points at N = 2
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/mgpu_check.py 2>&1 | grep -v "^registered" | tail -14 | tee gpurun_out/r2c16_mgpu_check.txt
: > gpurun_out/r2c16_scaling.txt
run() { n=$1; shift; out=$(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary "$@" 2>&1 | grep '^{' | tail -1); echo "N=$n $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['scaling'])" 2>/dev/null || echo FAILED)" | tee -a gpurun_out/r2c16_scaling.txt; }
run 2 --scaling strong
run 2
run 2 --policy FP64FP32
timeout 600 python -m pytest tests/test_native_slab_gpu.py -m gpu -q -p no:cacheprovider -rfEs 2>&1 | grep -v "^registered bc\|^$" | tail -5 | tee gpurun_out/r2c16_pytest_slab.txt
