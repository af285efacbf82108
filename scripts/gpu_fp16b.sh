mkdir -p gpurun_out; : > gpurun_out/fp16b.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest.log
run() { tag=$1; shift; if [ "$tag" = base ]; then unset XLB_B200_LIB; else export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_$tag.so; fi
  out=$(timeout 400 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$tag $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" 2>/dev/null || echo "FAILED: $out" | cut -c1-300)" | tee -a gpurun_out/fp16b.log; }
run base --policy FP32FP16
run base --policy FP32FP16 --config periodic
run base --policy FP32FP16 --lattice D3Q27
run base --policy FP32FP16 --lattice D3Q27 --config periodic
run q27_6 --policy FP32FP16 --lattice D3Q27 --config periodic
run q27_8 --policy FP32FP16 --lattice D3Q27 --config periodic
run base --policy FP64FP32
run q19_8 --policy FP64FP32
run q19_9 --policy FP64FP32
run base
run base --policy FP32FP16 --n 256
