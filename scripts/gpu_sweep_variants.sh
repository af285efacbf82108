# usage: bash scripts/gpu_sweep_variants.sh "<extra bench args>" tag1 tag2 ...   (tags = xlb_b200/variants/libxlb_b200_<tag>.so; "base" = product lib)
extra="$1"; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = base ]; then unset XLB_B200_LIB; else export XLB_B200_LIB=$PWD/xlb_b200/variants/libxlb_b200_$tag.so; fi
  for cfg in cavity periodic; do
    out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --config $cfg $extra 2>&1 | tail -1)
    echo "$tag $cfg $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-300)" | tee -a gpurun_out/sweep.log
  done
done
