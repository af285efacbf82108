mkdir -p gpurun_out; : > gpurun_out/sweep2.log
run() { tag=$1; shift; if [ "$tag" = base ]; then unset XLB_B200_LIB; else export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_$tag.so; fi
  out=$(timeout 400 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$tag $* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" 2>/dev/null || echo "FAILED: $out" | cut -c1-300)" | tee -a gpurun_out/sweep2.log; }
for t in kbc_0 kbc_5 kbc_8; do run $t --lattice D3Q27 --collision KBC --cells-per-thread 1; done
run kbc_5 --lattice D3Q27 --collision KBC --policy FP32FP16 --cells-per-thread 1
for t in q19_0 q19_5; do run $t --policy FP32FP16 --cells-per-thread 1; run $t --policy FP32FP16 --cells-per-thread 102; done
run q19_0 --cells-per-thread 1
run q19_0 --policy FP32FP16 --cells-per-thread 1 --config periodic
export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_kbc_5.so; python -m pytest tests/test_native_step_gpu.py -q -m gpu -k "kbc or c3" 2>&1 | tail -2
export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_q19_0.so; python -m pytest tests/test_native_step_gpu.py -q -m gpu -k "d3q19" 2>&1 | tail -2
