# Round 2, GPU call 1:  gpurun --timeout 1500 -- 'bash scripts/gpu_r2_call1.sh'
# diagnosis of the six round-1 GPU failures + the A/B timings that decide round 2's defaults
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r2_gpu.txt
timeout 600 python scripts/gpu_diag_r2.py 2>&1 | tail -150 | tee gpurun_out/r2_diag.log
: > gpurun_out/r2_matrix.log
run() { out=$(timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2_matrix.log; }
run
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --collision KBC --cells-per-thread 301
run --lattice D3Q27 --collision KBC --config sphere
run --lattice D3Q27 --collision KBC --config sphere --cells-per-thread 301
run --lattice D3Q27 --collision KBC --policy FP32FP16
run --lattice D3Q27 --collision KBC --policy FP32FP16 --cells-per-thread 301
run --policy FP32FP16
run --policy FP32FP16 --cells-per-thread 203
run --policy FP32FP16 --config periodic
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27 --policy FP32FP16 --cells-per-thread 203
run --policy FP64FP32
run --collision SmagorinskyLESBGK
run --config periodic --force 1e-6
run --lattice D3Q27 --collision KBC --config periodic --force 1e-6
python - <<'PY' 2>&1 | tee gpurun_out/r2_small_grids.log
import sys, time, torch
sys.path.insert(0, ".")
import bench
for n in (64, 128, 256):
    sys.argv = ["bench.py", "--n", str(n)]
    args = bench.parse()
    grid, stepper = bench.build_case(args, (n, n, n))
    f_0, f_1, bm, mm = stepper.prepare_fields()
    steps = 2000
    for label, fn in (("loop ", None), ("graph", True)):
        torch.cuda.synchronize(); t = time.perf_counter()
        if fn is None:
            for i in range(steps):
                f_0, f_1 = stepper(f_0, f_1, bm, mm, 1.0, i); f_0, f_1 = f_1, f_0
        else:
            f_0, f_1 = stepper.run(f_0, f_1, bm, mm, 1.0, steps)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{n}^3 {label}: {n**3 * steps / dt / 1e9:.1f} GLUPS ({dt / steps * 1e6:.1f} us/step)")
PY
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r2_kbc_lean python bench.py --n 256 --lattice D3Q27 --collision KBC --cells-per-thread 301 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_lean.log 2>&1
ncu -i gpurun_out/r2_kbc_lean.ncu-rep --page raw --csv > gpurun_out/r2_kbc_lean_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_kbc_lean.ncu-rep --page source --csv > gpurun_out/r2_kbc_lean_source.csv 2>/dev/null; rm -f gpurun_out/r2_kbc_lean.ncu-rep
