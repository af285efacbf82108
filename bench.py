#!/usr/bin/env python
"""Benchmark of the fused lattice-Boltzmann step on B200 (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): MLUPS = cells * steps / seconds / 1e6 (reference: examples/performance/mlups_3d.py:87-90).
Workload at N=1 (BASELINE.json configs[1]): D3Q19 BGK 512^3 FP32FP32 lid-driven cavity, i.e. exactly the set-up of the
reference's own benchmark script (mlups_3d.py:45-63: EquilibriumBC lid u=(0.02,0,0), FullwayBounceBack walls, omega=1)
at the edge length BASELINE names.  N>1: weak scaling, 512^3 per GPU as x-slabs (4096x512x512 at 8 GPUs).

One "step" = one LBM time step = ONE launch of the fused kernel (plus 2 face-plane launches and 2 one-thread
signal/wait kernels per step on slab grids).  Inputs are 20 GB >> 126 MB L2, so no L2 flush is needed between steps.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STORE_BYTES = {"FP32FP32": 4, "FP64FP32": 4, "FP32FP16": 2, "FP64FP16": 2, "FP64FP64": 8}
Q = {"D3Q19": 19, "D3Q27": 27}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--n", type=int, default=512, help="edge length per GPU")
    p.add_argument("--lattice", default="D3Q19", choices=list(Q))
    p.add_argument("--collision", default="BGK", choices=["BGK", "KBC", "SmagorinskyLESBGK"])
    p.add_argument("--force", type=float, default=0.0, help="x-component of a constant body force (ForcedCollision / ExactDifference); 0 = none")
    p.add_argument("--policy", default="FP32FP32", choices=list(STORE_BYTES))
    p.add_argument("--config", default="cavity", choices=["cavity", "periodic", "sphere", "tunnel"])
    p.add_argument("--cells-per-thread", type=int, default=0)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=0)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in open(self.path):
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# workload through the public operator API (a reference-style script)
# ---------------------------------------------------------------------------------------------------------------------


def build_case(args, shape):
    import xlb_b200 as xlb
    from xlb_b200.compute_backend import ComputeBackend
    from xlb_b200.grid import grid_factory
    from xlb_b200.operator.boundary_condition import EquilibriumBC, FullwayBounceBackBC
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    be = ComputeBackend.WARP
    pp = xlb.PrecisionPolicy[args.policy]
    vs = getattr(xlb.velocity_set, args.lattice)(precision_policy=pp, compute_backend=be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    grid = grid_factory(shape)
    bcs = []
    if args.config == "cavity":  # examples/performance/mlups_3d.py:45-63
        box = grid.bounding_box_indices()
        box_no_edge = grid.bounding_box_indices(remove_edges=True)
        lid = box_no_edge["top"]
        walls = [box["bottom"][i] + box["left"][i] + box["right"][i] + box["front"][i] + box["back"][i] for i in range(3)]
        walls = np.unique(np.array(walls), axis=-1).tolist()
        bcs = [EquilibriumBC(rho=1.0, u=(0.02, 0.0, 0.0), indices=lid), FullwayBounceBackBC(indices=walls)]
    elif args.config in ("sphere", "tunnel"):
        bcs = obstacle_bcs(args.config, grid, shape)
    kw = dict(force_vector=np.array([args.force, 0.0, 0.0])) if args.force else {}
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=args.collision, cells_per_thread=args.cells_per_thread, **kw)
    return grid, stepper


def obstacle_bcs(kind, grid, shape):
    """C3: flow past a sphere (examples/cfd/flow_past_sphere_3d.py:41-109 geometry: Fullway walls, Regularized velocity
    inlet with a Poiseuille profile, ExtrapolationOutflow outlet, Halfway sphere by index inequality).
    C4: the same tunnel with a synthetic voxelised bluff body (seeded union of boxes + an ellipsoid, seed 0) and a uniform
    inlet (examples/cfd/windtunnel_3d.py:92-96 boundary set)."""
    from xlb_b200.operator.boundary_condition import ExtrapolationOutflowBC, FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC

    nx, ny, nz = shape
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    if kind == "sphere":
        r = ny // 12
        cx, cy, cz = nx // 6, ny // 2, nz // 2
        xs = np.arange(cx - r, cx + r + 1)
        X, Y, Z = np.meshgrid(xs, np.arange(cy - r, cy + r + 1), np.arange(cz - r, cz + r + 1), indexing="ij")
        m = (X - cx) ** 2 + (Y - cy) ** 2 + (Z - cz) ** 2 < r**2
        body = [X[m].tolist(), Y[m].tolist(), Z[m].tolist()]
        Hy, Hz = float(ny - 1), float(nz - 1)

        def profile(index):
            yc, zc = index[1] - Hy / 2.0, index[2] - Hz / 2.0
            return [0.04 * np.maximum(0.0, 1.0 - ((2.0 * yc / Hy) ** 2.0 + (2.0 * zc / Hz) ** 2.0))]

        inlet = RegularizedBC("velocity", profile=profile, indices=bne["left"])
    else:
        rng = np.random.default_rng(0)
        solid = np.zeros((nx // 4, ny // 2, nz // 3), dtype=bool)  # body region x in [nx/4, nx/2)
        sx, sy, sz = solid.shape
        for _ in range(6):
            lo = [rng.integers(0, s // 2) for s in solid.shape]
            hi = [l + rng.integers(s // 4, s // 2) for l, s in zip(lo, solid.shape)]
            solid[lo[0] : hi[0], lo[1] : hi[1], lo[2] : hi[2]] = True
        X, Y, Z = np.meshgrid(np.arange(sx), np.arange(sy), np.arange(sz), indexing="ij")
        solid |= ((X - sx / 2) / (sx / 2.2)) ** 2 + ((Y - sy / 2) / (sy / 2.5)) ** 2 + ((Z - sz / 3) / (sz / 3.0)) ** 2 < 1.0
        solid[:, :, :2] = False
        ix, iy, iz = np.nonzero(solid)
        body = [(ix + nx // 4).tolist(), (iy + ny // 4).tolist(), (iz + 2).tolist()]
        inlet = RegularizedBC("velocity", prescribed_value=(0.02, 0.0, 0.0), indices=bne["left"])
    return [FullwayBounceBackBC(indices=walls), inlet, ExtrapolationOutflowBC(indices=bne["right"]), HalfwayBounceBackBC(indices=body)]


def run_native(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:  # launched without torchrun: re-exec under it
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                   "--master-port", "29511"] + sys.argv  # fmt: skip
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    shape = (n * world, n, n)  # weak scaling: n^3 per GPU, x-slabs
    if args.config in ("sphere", "tunnel"):  # BASELINE C3: 1024x512x512 (= 2n x n x n); C4: 1152x512x512 over 8 GPUs (= 9n/4 x n x n)
        shape = (2 * n, n, n) if args.config == "sphere" else (9 * n // 4 // world * world, n, n)
    grid, stepper = build_case(args, shape)
    f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()
    omega = 1.0 if args.config in ("cavity", "periodic") else 1.6
    cells_total = shape[0] * shape[1] * shape[2]
    cells_local = cells_total // world

    def loop(k, t0):
        nonlocal f_0, f_1
        for i in range(k):
            f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, t0 + i)
            f_0, f_1 = f_1, f_0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loop(max(args.warmup, 3), 0)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    loop(args.steps, args.warmup)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    mlups = cells_total * args.steps / (ms * 1e-3) / 1e6
    finite = bool(torch.isfinite(f_0[:, :: max(1, n // 8)]).all())

    # roofline of the dominant kernel: one fused-step launch per step covers the slab (interior launch on slab grids)
    bytes_per_cell = 2 * Q[args.lattice] * STORE_BYTES[args.policy] + 1
    peak, peak_src = peaks()
    achieved = bytes_per_cell * cells_local / (ms_per_step * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": None, "bytes_per_cell": bytes_per_cell, "peak_source": peak_src, "kernel": "xlbn::step_kernel",
        "launch_ms": round(ms_per_step, 4),
    }  # fmt: skip
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            key = f"{args.lattice}_{args.collision}_{args.policy}_{n}"
            roofline["traffic"] = json.load(open(traffic_file)).get(key)
        except Exception:
            pass

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, stepper, grid, (f_0, f_1, bc_mask, missing_mask), cells_total, world, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    launches_per_step = 1 if world == 1 else 3 + 2  # interior + 2 face planes + wait + signal
    if rank == 0:
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak" if args.config in ("cavity", "periodic") else "strong", "vs_baseline": None,
            "dtype": {"FP32FP32": "f32", "FP32FP16": "f32 compute / f16 store", "FP64FP32": "f64 compute / f32 store", "FP64FP64": "f64", "FP64FP16": "f64 compute / f16 store"}[args.policy],
            "data": "synthetic",
            "config": {
                "workload": f"{args.config} {args.lattice} {args.collision} {n}^3 per GPU {args.policy} (global {shape[0]}x{shape[1]}x{shape[2]}); "
                            + {"cavity": "lid-driven cavity of examples/performance/mlups_3d.py", "periodic": "fully periodic box",
                               "sphere": "flow past a sphere (examples/cfd/flow_past_sphere_3d.py geometry)",
                               "tunnel": "synthetic voxelised bluff body in a wind tunnel (windtunnel_3d.py boundary set)"}[args.config],
                "omega": omega, "l2": "inputs (2 x %.1f GB per GPU) larger than L2, no flush" % (Q[args.lattice] * STORE_BYTES[args.policy] * cells_local / 1e9),
                "parallelism": "1 GPU" if world == 1 else f"x-slab x{world}, halo fused into the face-plane kernels (peer stores over NVLink)",
                "cells_per_thread": args.cells_per_thread or "default",
            },
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "finite": finite,
        }  # fmt: skip
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, stepper, grid, fields, cells_total, world, barrier):
    """End to end through the public operator API with HOST buffers: the job's populations and masks start in pinned host
    memory, are copied to the device inside the timed region, stepped K_e times with `stepper(...)` (per step the host
    also sends `omega` and reads back a monitoring probe: the mid-x plane of all populations), and the final populations
    are copied back to pinned host memory.  MLUPS = cells * K_e / wall-clock-on-device of all of that."""
    import torch

    f_0, f_1, bc_mask, missing_mask = fields
    k = args.e2e_steps or args.steps
    nx = f_0.shape[1]
    try:
        h_f = torch.empty(f_0.shape, dtype=f_0.dtype, pin_memory=True)
        h_out = torch.empty(f_0.shape, dtype=f_0.dtype, pin_memory=True)
        h_bc = torch.empty(bc_mask.shape, dtype=bc_mask.dtype, pin_memory=True)
        h_mm = torch.empty(missing_mask.shape, dtype=missing_mask.dtype, pin_memory=True)
        h_probe = torch.empty((f_0.shape[0],) + tuple(f_0.shape[2:]), dtype=f_0.dtype, pin_memory=True)
        h_omega = torch.ones(1, dtype=torch.float64, pin_memory=True)
    except RuntimeError as e:
        return {"value": None, "unit": "MLUPS", "error": f"pinned allocation failed: {e}"}
    h_f.copy_(f_0)
    h_bc.copy_(bc_mask)
    h_mm.copy_(missing_mask)
    d_omega = torch.empty(1, dtype=torch.float64, device=f_0.device)
    stepper.reset_halo() if world > 1 else None
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    f_0.copy_(h_f, non_blocking=True)
    f_1.copy_(f_0)
    bc_mask.copy_(h_bc, non_blocking=True)
    missing_mask.copy_(h_mm, non_blocking=True)
    for i in range(k):
        d_omega.copy_(h_omega, non_blocking=True)
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, 1.0, i)
        f_0, f_1 = f_1, f_0
        h_probe.copy_(f_0[:, nx // 2], non_blocking=True)
    h_out.copy_(f_0, non_blocking=True)
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h2d = (h_f.numel() * h_f.element_size() + h_bc.numel() + h_mm.numel()) / k + 8
    d2h = (h_out.numel() * h_out.element_size()) / k + h_probe.numel() * h_probe.element_size()
    return {
        "value": round(cells_total * k / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
        "steps": k, "what": "pinned-host populations+masks -> device, K steps via stepper(...), per-step probe plane D2H, final populations -> pinned host",
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's restatement of the reference's JAX path on the host cores
# ---------------------------------------------------------------------------------------------------------------------


def cpu_lbm(args, n, steps, threads):
    """Time `steps` steps of the same workload at edge n on the CPU; returns (MLUPS, kind, cores)."""
    sys.path.insert(0, ROOT)
    try:
        from oracle import lbm_c

        if lbm_c.available():
            mlups = lbm_c.time_cavity(args.lattice, args.collision, args.policy, n, steps, threads, periodic=(args.config == "periodic"))
            return mlups, "port", threads
    except ImportError:
        pass
    from oracle import lbm_numpy as O

    lat = O.Lattice(args.lattice)
    shape = (n, n, n)
    bcs = []
    if args.config == "cavity":
        box, box_ne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
        walls = np.unique(np.concatenate([box[k] for k in ("bottom", "left", "right", "front", "back")], axis=1), axis=-1)
        bcs = [O.BC("equilibrium", 1, box_ne["top"], rho=1.0, u=(0.02, 0.0, 0.0)), O.BC("fullway", 2, walls)]
        bc_mask, missing = O.build_masks(bcs, shape, lat, flavor="warp")
    else:
        bc_mask, missing = np.zeros((1,) + shape, np.uint8), np.zeros((lat.q,) + shape, bool)
    f = O.initialize_eq(shape, lat, args.policy)
    f = O.run(f, bc_mask, missing, bcs, 1.0, lat, 1, policy=args.policy, collision=args.collision)
    t0 = time.perf_counter()
    O.run(f, bc_mask, missing, bcs, 1.0, lat, steps, policy=args.policy, collision=args.collision)
    dt = time.perf_counter() - t0
    return n**3 * steps / dt / 1e6, "port", 1


def cpu_baseline(args):
    threads = os.cpu_count() or 1
    n, steps = 128, 10
    mlups, kind, cores = cpu_lbm(args, n, steps, threads)
    return {
        "value": round(mlups, 2), "unit": "MLUPS", "cores": cores, "kind": kind,
        "sample": f"{args.config} {args.lattice} {args.collision} {args.policy} at {n}^3 for {steps} steps (bounded sample of the 512^3 workload; "
                  "CPU restatement of the reference's step, the reference itself needs jax/warp which are not installed)",
    }  # fmt: skip


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 128
    per_step = 4  # LBM steps per bench "step": a bounded sample of the workload
    cpu_lbm(args, n, 1, threads)
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, args.warmup)):
        cpu_lbm(args, n, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, kind, cores = cpu_lbm(args, n, per_step, threads)
        vals.append(v)
        if time.perf_counter() - t0 > 150:
            break
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = f"{args.config} {args.lattice} {args.collision} {args.policy} at {n}^3, {per_step} LBM steps per bench step, {len(vals)} bench steps"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(value, 2), "unit": "MLUPS", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": round(elapsed / max(1, len(vals)) * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} {args.lattice} {args.collision} 512^3 {args.policy} (timed on a bounded {n}^3 sample)"},
        "cpu_baseline": {"value": round(value, 2), "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's CPU path: jax / warp are not installed in this image, so this is the oracle's restatement of the reference's step on the host cores",
    }  # fmt: skip
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
