# First GPU call of round 2 (one B200):   gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh'
# 1. the validated suite, 2. first execution of everything written after round 1's GPU budget was spent,
# 3. A/B timings that decide round 2's defaults (lean KBC vs default KBC) and first numbers for the extended collision kernels.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_zz_first_run_gpu.py 2>&1 | tail -6 | tee gpurun_out/r2_pytest_validated.log
# -rxX: list xfailed (= a late case FAILED on its first run: read the reason) and xpassed (= it works: promote it)
timeout 1500 python -m pytest tests/test_zz_first_run_gpu.py -m gpu -q -rxX -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/r2_pytest_first_run.log
: > gpurun_out/r2_matrix.log
run() { out=$(timeout 400 python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1); echo "$* => $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'].get('sm_mhz'))" 2>/dev/null || echo "FAILED: $out" | cut -c1-400)" | tee -a gpurun_out/r2_matrix.log; }
run                                                              # headline, must still read ~41 GLUPS / 0.97
# KBC: default formulation vs the register-lean one (cells_per_thread 301)
run --lattice D3Q27 --collision KBC
run --lattice D3Q27 --collision KBC --cells-per-thread 301
run --lattice D3Q27 --collision KBC --config periodic
run --lattice D3Q27 --collision KBC --config periodic --cells-per-thread 301
run --lattice D3Q27 --collision KBC --config sphere
run --lattice D3Q27 --collision KBC --config sphere --cells-per-thread 301
run --lattice D3Q27 --collision KBC --policy FP32FP16
run --lattice D3Q27 --collision KBC --policy FP32FP16 --cells-per-thread 301
# wind tunnel with a mesh body (N3): voxelisation + 200 steps
timeout 300 python examples/windtunnel_mesh.py 256 96 96 200 2>&1 | tail -14 | tee gpurun_out/r2_windtunnel_mesh.log
timeout 300 python examples/turbulent_channel.py 32 400 2>&1 | tail -8 | tee gpurun_out/r2_turbulent_channel.log
# FP32FP16: half2-state path (default, 202) vs the split boundary variant (203)
run --policy FP32FP16
run --policy FP32FP16 --cells-per-thread 203
run --policy FP32FP16 --config periodic
run --lattice D3Q27 --policy FP32FP16
run --lattice D3Q27 --policy FP32FP16 --cells-per-thread 203
# extended collision kernels (N4): first numbers
run --collision SmagorinskyLESBGK
run --collision SmagorinskyLESBGK --config periodic
run --config periodic --force 1e-6
run --collision SmagorinskyLESBGK --config periodic --force 1e-6
run --lattice D3Q27 --collision SmagorinskyLESBGK --config periodic
run --lattice D3Q27 --collision KBC --config periodic --force 1e-6
run --collision SmagorinskyLESBGK --policy FP32FP16
run --config periodic --force 1e-6 --policy FP64FP64 --n 384
# small grids: the user loop vs stepper.run (CUDA graph of a pair of steps)
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
# one full capture of the lean KBC kernel for the register / stall picture
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r2_kbc_lean python bench.py --n 256 --lattice D3Q27 --collision KBC --cells-per-thread 301 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_lean.log 2>&1
ncu -i gpurun_out/r2_kbc_lean.ncu-rep --page raw --csv > gpurun_out/r2_kbc_lean_raw.csv 2>/dev/null && rm -f gpurun_out/r2_kbc_lean.ncu-rep
