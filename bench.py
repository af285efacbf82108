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
    p.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="N > 1: weak = n^3 per GPU (x-slabs), strong = the N = 1 grid split over N GPUs")
    p.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (FP32FP16 cavity, D3Q27 KBC cavity, C3 sphere) appended at N = 1")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=0)
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlClockSampler:
    """SM clock and clock-event (throttle) reasons through NVML — what nvidia-smi itself reads — sampled every 5 ms by a thread of this
    process, time-stamped, so that even a 60 ms timed region holds a dozen samples.  Preferred over the nvidia-smi child process below,
    whose first line can arrive after a short region has ended (and whose buffered output is lost when it is terminated)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, torch, device_index):
        import threading

        import pynvml

        pynvml.nvmlInit()
        self.nv = pynvml
        handle = None
        try:  # the CUDA ordinal is not the NVML index under CUDA_VISIBLE_DEVICES: go through the UUID
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            handle = None
        self.handle = handle if handle is not None else pynvml.nvmlDeviceGetHandleByIndex(device_index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        self.read_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.sample()  # fails here, not in the thread, if the queries are unsupported
        self.rows, self.stop_flag = [], threading.Event()
        self.thread = threading.Thread(target=self.loop, daemon=True)

    def sample(self):
        return (time.time(), float(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)), int(self.read_reasons(self.handle)))

    def loop(self):
        while not self.stop_flag.is_set():
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            self.stop_flag.wait(0.005)

    def start(self):
        self.thread.start()

    def stop(self, t_begin=None, t_end=None):
        self.stop_flag.set()
        self.thread.join(timeout=2)
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
        inside = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (no sample fell inside the region)"
        bits = 0
        for r in inside:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": self.max_mhz, "reasons": [n for n, b in self.REASONS if bits & b],
                "samples": len(inside), "window": window, "source": "NVML (pynvml), 5 ms period"}  # fmt: skip


def clock_sampler(torch, device_index):
    try:
        return NvmlClockSampler(torch, device_index)
    except Exception:
        return ClockSampler(device_index)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Started before the warm-up (the
    process takes ~0.1 s to produce its first line, longer than a 20-step timed region at 512^3); samples are time-stamped and
    the ones inside [t_begin, t_end] of the timed region are used, falling back to the samples taken since the warm-up began
    (the GPU is under the same load there) when the region was shorter than one sampling period."""

    FIELDS = "timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """t_begin / t_end: time.time() bracketing the timed region."""
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[2]), float(parts[3]), parts[6:10]))
            except ValueError:
                continue
        os.unlink(self.path)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        window = "timed region"
        if not inside:
            inside, window = rows, "warm-up + timed region (the region was shorter than one sampling period)"
        reasons = set()
        for r in inside:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)), "reasons": sorted(reasons),
                "samples": len(inside), "window": window}  # fmt: skip


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


DTYPE_NAME = {"FP32FP32": "f32", "FP32FP16": "f32 compute / f16 store", "FP64FP32": "f64 compute / f32 store", "FP64FP64": "f64", "FP64FP16": "f64 compute / f16 store"}
WHAT = {"cavity": "lid-driven cavity of examples/performance/mlups_3d.py", "periodic": "fully periodic box",
        "sphere": "flow past a sphere (examples/cfd/flow_past_sphere_3d.py geometry)",
        "tunnel": "synthetic voxelised bluff body in a wind tunnel (windtunnel_3d.py boundary set)"}  # fmt: skip


def global_shape(args, world):
    n = args.n
    if args.config == "sphere":  # BASELINE C3: 1024x512x512 (= 2n x n x n)
        return (2 * n, n, n)
    if args.config == "tunnel":  # BASELINE C4: 1152x512x512 over 8 GPUs (= 9n/4 x n x n, rounded to the slab count)
        return (9 * n // 4 // world * world, n, n)
    if args.scaling == "strong":
        return (n, n, n)
    return (n * world, n, n)  # weak scaling: n^3 per GPU, x-slabs


def scaling_of(args):
    return "strong" if (args.scaling == "strong" or args.config in ("sphere", "tunnel")) else "weak"


def config_dict(args, world, shape, omega):
    """The `config` object of the JSON line; the native and the reference arm print the same one for the same arguments."""
    per_gpu = f"{args.n}^3 per GPU" if scaling_of(args) == "weak" else f"{shape[0] // world}x{shape[1]}x{shape[2]} per GPU"
    return {
        "workload": f"{args.config} {args.lattice} {args.collision} {per_gpu} {args.policy} (global {shape[0]}x{shape[1]}x{shape[2]}); {WHAT[args.config]}",
        "omega": omega, "n_gpus": world,
        "l2": "inputs (2 x %.1f GB per GPU) larger than L2, no flush" % (Q[args.lattice] * STORE_BYTES[args.policy] * shape[0] * shape[1] * shape[2] / world / 1e9),
        "parallelism": "1 GPU" if world == 1 else f"x-slab x{world}, halo fused into the face-plane kernels (peer stores over NVLink)",
        "cells_per_thread": args.cells_per_thread or "default",
    }  # fmt: skip


def kernel_label(args, shape):
    """Name of the kernel the library picks for this workload (csrc/step_kernel.cuh: launch_step / launch_step_base), for the roofline
    object; an ncu capture of the same workload (profiles/traffic.json) overrides it with the name the profiler saw."""
    ny, nz = shape[1], shape[2]
    tiles = lambda cells, mult: nz <= cells and cells % nz == 0 and nz % mult == 0 and ny % (cells // nz) == 0  # noqa: E731
    v = int(getattr(args, "cells_per_thread", 0) or 0)
    forced = bool(getattr(args, "force", 0.0))
    if args.policy == "FP32FP16" and args.collision == "BGK" and not forced and v in (0, 402, 404) and (tiles(1024, 8) or tiles(512, 8)):
        return "xlbn::step_tile_kernel (TMA-fed, half2 pairs)"
    if args.lattice == "D3Q19" and args.policy in ("FP32FP32", "FP64FP32") and tiles(512, 16) and v in (0, 501, 502):
        if args.collision == "BGK" and not forced or v == 501 or (args.collision == "SmagorinskyLESBGK" and not forced and args.policy == "FP32FP32"):
            return "xlbn::step_tile1_kernel (TMA-fed, one cell per thread)"
    return "xlbn::step_kernel (direct loads)"


def roofline_of(args, shape, world, ms_per_step):
    """Roofline of the dominant kernel: ONE fused-step launch per step covers the slab (the interior launch on slab grids)."""
    bytes_per_cell = 2 * Q[args.lattice] * STORE_BYTES[args.policy] + 1
    cells_local = shape[0] * shape[1] * shape[2] // world
    peak, peak_src = peaks()
    achieved = bytes_per_cell * cells_local / (ms_per_step * 1e-3) / 1e9
    out = {
        "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": None, "bytes_per_cell": bytes_per_cell, "bytes_per_launch": bytes_per_cell * cells_local, "peak_source": peak_src,
        "kernel": kernel_label(args, (shape[0] // world, shape[1], shape[2])), "launch_ms": round(ms_per_step, 4),
    }  # fmt: skip
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):  # ncu --set full captures, keyed by the complete workload (config, lattice, collision, policy, local extents)
        try:
            key = f"{args.config}_{args.lattice}_{args.collision}{'_forced' if getattr(args, 'force', 0.0) else ''}_{args.policy}_{shape[0] // world}x{shape[1]}x{shape[2]}"
            if str(getattr(args, "cells_per_thread", 0)) not in ("0", "default"):
                key += f"_v{args.cells_per_thread}"  # a capture belongs to the kernel the library picks by default
            entry = json.load(open(traffic_file)).get(key)
            if entry is not None:
                out["traffic"] = entry["bytes"] if isinstance(entry, dict) else entry
                if isinstance(entry, dict):
                    out["traffic_source"] = entry.get("source")
                    if entry.get("kernel"):  # the kernel the capture saw for this workload (the library picks it: csrc/step_kernel.cuh launch_step_base)
                        out["kernel"] = "xlbn::" + entry["kernel"].replace("void ", "").strip()
        except Exception:
            pass
    return out


def timed_steps(torch, stepper, fields, omega, steps, warmup, barrier, t0=0):
    """W untimed steps, then exactly K steps between CUDA events on the current stream.  Returns (ms, fields, (t_begin, t_end))."""
    f_0, f_1, bc_mask, missing_mask = fields
    for i in range(warmup):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, t0 + i)
        f_0, f_1 = f_1, f_0
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    start.record()
    for i in range(steps):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, t0 + warmup + i)
        f_0, f_1 = f_1, f_0
    stop.record()
    barrier()
    return start.elapsed_time(stop), (f_0, f_1, bc_mask, missing_mask), (w0, time.time())


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

    shape = global_shape(args, world)
    grid, stepper = build_case(args, shape)
    fields = stepper.prepare_fields()
    omega = 1.0 if args.config in ("cavity", "periodic") else 1.6
    cells_total = shape[0] * shape[1] * shape[2]
    warmup = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = clock_sampler(torch, local_rank) if rank == 0 else None
    if rank == 0:
        sampler.start()
    ms, fields, window = timed_steps(torch, stepper, fields, omega, args.steps, warmup, barrier)
    clocks = sampler.stop(*window) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    mlups = cells_total * args.steps / (ms * 1e-3) / 1e6
    finite = bool(torch.isfinite(fields[0][:, :: max(1, args.n // 8)]).all())
    roofline = roofline_of(args, shape, world, ms_per_step)

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, stepper, grid, fields, cells_total, world, barrier)

    secondary = None
    if world == 1 and not args.no_secondary and (args.config, args.lattice, args.collision, args.policy, args.n) == ("cavity", "D3Q19", "BGK", "FP32FP32", 512):
        del fields, stepper, grid
        secondary = run_secondary(args, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    launches_per_step = 1 if world == 1 else 3 + 2  # interior + 2 face planes + wait + signal
    if rank == 0:
        line = {
            "metric": "MLUPS", "value": round(mlups, 1), "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None,
            "dtype": DTYPE_NAME[args.policy], "data": "synthetic", "config": config_dict(args, world, shape, omega),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "finite": finite,
        }  # fmt: skip
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args, barrier):
    """The other BASELINE workloads that fit one GPU, timed the same way (20 steps each, CUDA events): so that the driver's record
    carries more than the headline configuration.  Each entry has its own roofline against the same measured peak."""
    import copy
    import gc

    import torch

    out = []
    for over in (dict(policy="FP32FP16"), dict(policy="FP32FP16", config="periodic"), dict(lattice="D3Q27", collision="KBC"),
                 dict(lattice="D3Q27", collision="KBC", config="sphere"), dict(lattice="D3Q27", policy="FP32FP16")):  # fmt: skip
        a = copy.copy(args)
        for k, v in over.items():
            setattr(a, k, v)
        try:
            gc.collect()
            torch.cuda.empty_cache()
            shape = global_shape(a, 1)
            grid, stepper = build_case(a, shape)
            fields = stepper.prepare_fields()
            omega = 1.0 if a.config in ("cavity", "periodic") else 1.6
            ms, fields, _ = timed_steps(torch, stepper, fields, omega, 20, 3, barrier)
            cells = shape[0] * shape[1] * shape[2]
            out.append({"workload": config_dict(a, 1, shape, omega)["workload"], "value": round(cells * 20 / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS",
                        "steps": 20, "warmup": 3, "ms_per_step": round(ms / 20, 4), "dtype": DTYPE_NAME[a.policy], "roofline": roofline_of(a, shape, 1, ms / 20),
                        "finite": bool(torch.isfinite(fields[0][:, :: max(1, a.n // 8)]).all())})  # fmt: skip
            del fields, stepper, grid
        except Exception as e:  # a secondary workload must never cost the headline line
            out.append({"workload": str(over), "error": f"{type(e).__name__}: {e}"[:300]})
    return out


def run_e2e(args, stepper, grid, fields, cells_total, world, barrier):
    """End to end through the public operator API with HOST buffers: the job's populations and masks start in pinned host memory and its
    final populations end there; MLUPS = cells * K_e / device-timed duration of ALL of that.

    N = 1: `stepper.run_streamed` — upload, K_e steps and download pipelined as a wavefront over x-planes (chunks of 32 planes; the
    stepper's own kernels on partial x ranges; bit-identical to K_e ordinary calls, tests/test_native_step_more_gpu.py).
    N > 1 (slab grids): the serial form — upload, K_e calls of `stepper(...)` with a per-step omega upload and probe-plane read-back,
    download."""
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
    streamed = world == 1 and 2 * k + 2 <= nx and hasattr(stepper, "run_streamed")
    d_omega = torch.empty(1, dtype=torch.float64, device=f_0.device)
    stepper.reset_halo() if world > 1 else None
    if streamed:  # one untimed pass: stream creation, first-use allocations
        stepper.run_streamed(h_f, h_out, f_0, f_1, bc_mask, missing_mask, 1.0, 2, host_bc_mask=h_bc, host_missing_mask=h_mm)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    if streamed:
        d_omega.copy_(h_omega, non_blocking=True)
        stepper.run_streamed(h_f, h_out, f_0, f_1, bc_mask, missing_mask, 1.0, k, host_bc_mask=h_bc, host_missing_mask=h_mm)
        mm_bytes = h_mm.numel() if stepper._needs_missing else 0
        h2d = (h_f.numel() * h_f.element_size() + h_bc.numel() + mm_bytes + 8) / k
        d2h = (h_out.numel() * h_out.element_size()) / k
        what = ("stepper.run_streamed: pinned-host populations + bc_mask" + (" + missing_mask" if mm_bytes else " (no boundary condition of this stepper reads missing_mask: not uploaded)")
                + " -> device in 32-plane chunks, K steps as a wavefront behind the upload, final populations -> pinned host behind the last step; all inside the timed region")
    else:
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
        h2d = (h_f.numel() * h_f.element_size() + h_bc.numel() + h_mm.numel()) / k + 8
        d2h = (h_out.numel() * h_out.element_size()) / k + h_probe.numel() * h_probe.element_size()
        what = "pinned-host populations+masks -> device, K steps via stepper(...), per-step probe plane D2H, final populations -> pinned host"
    stop.record()
    barrier()
    ms = start.elapsed_time(stop)
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {
        "value": round(cells_total * k / (ms * 1e-3) / 1e6, 1), "unit": "MLUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
        "steps": k, "ms": round(ms, 2), "finite": bool(torch.isfinite(h_out[:, :: max(1, nx // 8)]).all()),
        "bytes_are": "per rank (each of the %d ranks moves this much per step)" % world, "what": what,
    }  # fmt: skip


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's restatement of the reference's JAX path on the host cores
# ---------------------------------------------------------------------------------------------------------------------


def cpu_runner(args, n, threads):
    """The oracle's C/OpenMP restatement of the reference step on the same workload at edge n, ready to be stepped (persistent
    host buffers: no copies inside the timed calls).  The ONLY use of oracle/ in this file besides nothing: the CPU legs."""
    from oracle import lbm_c

    if not lbm_c.available():
        raise RuntimeError("oracle/liblbm_ref.so is not built (python -c 'import __graft_entry__ as g; g.build()')")
    return lbm_c.cavity_runner(args.lattice, args.collision, args.policy, n, threads, periodic=(args.config == "periodic"))


def cpu_baseline(args):
    """Bounded sample for the native arm's line: the same workload at 128^3, 3 x 10 steps after a warm-up step, all host threads."""
    threads = os.cpu_count() or 1
    n, steps, reps = 128, 10, 3
    r = cpu_runner(args, n, threads)
    r.steps(1)
    vals = [n**3 * steps / r.steps(steps) / 1e6 for _ in range(reps)]
    return {
        "value": round(float(np.mean(vals)), 2), "unit": "MLUPS", "cores": threads, "kind": "port",
        "sample": f"{args.config} {args.lattice} {args.collision} {args.policy} at {n}^3, mean of {reps} x {steps} steps (bounded sample of the 512^3 workload; the "
                  "reference arm, bench.py --impl reference, times the full-size grid; CPU restatement of the reference's step — the reference itself needs jax / warp, not installed)",
    }  # fmt: skip


def run_reference(args):
    """Reference arm: the reference's CPU path for the SAME configuration (512^3 by default: 22 GB of host memory), one LBM step per
    bench step, all host threads.  jax / warp are not installable here, so the stepping code is the oracle's C/OpenMP
    restatement of the reference's fused Warp kernel, which reproduces the reference's own WARP backend bit for bit
    (tests/test_warp_path_golden.py).  Stops after ~200 s of stepping and reports the steps it completed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config not in ("cavity", "periodic"):
        raise SystemExit("--impl reference: cavity / periodic workloads")
    threads = os.cpu_count() or 1
    world = args.gpus
    shape = global_shape(args, world)
    n = args.n
    try:
        r = cpu_runner(args, n, threads)
        sample = f"the full {n}^3 grid of one rank, 1 LBM step per bench step"
    except MemoryError:
        n = 256
        r = cpu_runner(args, n, threads)
        sample = f"{n}^3 (the host could not hold the full {args.n}^3 grid), 1 LBM step per bench step"
    budget, t_all = 200.0, time.perf_counter()
    for _ in range(min(max(args.warmup, 1), 2)):
        r.steps(1)
    times = []
    for _ in range(args.steps):
        times.append(r.steps(1))
        if time.perf_counter() - t_all > budget:
            break
    value = n**3 * len(times) / sum(times) / 1e6
    if len(times) < args.steps:
        sample += f"; stopped after {len(times)} of {args.steps} steps ({budget:.0f} s budget)"
    line = {
        "impl": "reference", "metric": "MLUPS", "value": round(value, 2), "unit": "MLUPS", "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": round(sum(times) / len(times) * 1e3, 2), "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None,
        "dtype": DTYPE_NAME[args.policy], "data": "synthetic", "config": config_dict(args, world, shape, 1.0),
        "cpu_baseline": {"value": round(value, 2), "unit": "MLUPS", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference's CPU path: jax / warp are not installed in this image, so this is the oracle's C/OpenMP restatement of the reference's fused step "
                "(bit-identical to the reference's WARP backend on the test vectors) on the host cores; one rank's grid, N ranks would each do the same",
    }  # fmt: skip
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
