"""Multi-GPU check, run under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/mgpu_check.py
Runs reference-style cases on an x-slab grid (fused peer-store halo, overlapped face/interior launches) and checks that
the gathered result is BIT-IDENTICAL to the same case run on one GPU without decomposition."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import xlb_b200 as xlb  # noqa: E402
from xlb_b200.compute_backend import ComputeBackend  # noqa: E402
from xlb_b200.grid import grid_factory  # noqa: E402
from xlb_b200.operator.boundary_condition import (  # noqa: E402
    EquilibriumBC,
    ExtrapolationOutflowBC,
    FullwayBounceBackBC,
    HalfwayBounceBackBC,
    RegularizedBC,
)
from xlb_b200.operator.boundary_condition.boundary_condition_registry import boundary_condition_registry as reg  # noqa: E402
from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper  # noqa: E402


def build(case, shape, lattice, policy, collision, distributed):
    reg.next_id = 1
    be = ComputeBackend.WARP
    pp = xlb.PrecisionPolicy[policy]
    vs = getattr(xlb.velocity_set, lattice)(precision_policy=pp, compute_backend=be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    grid = grid_factory(shape, distributed=distributed)
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    if case == "cavity2d":  # examples/cfd/lid_driven_cavity_2d(_distributed).py: Halfway walls, EquilibriumBC lid
        walls = [box["bottom"][i] + box["left"][i] + box["right"][i] for i in range(2)]
        walls = np.unique(np.array(walls), axis=-1).tolist()
        bcs = [EquilibriumBC(rho=1.0, u=(0.05, 0.0), indices=bne["top"]), HalfwayBounceBackBC(indices=walls)]
        omega = 1.5
    elif case == "cavity":
        walls = [box["bottom"][i] + box["left"][i] + box["right"][i] + box["front"][i] + box["back"][i] for i in range(3)]
        walls = np.unique(np.array(walls), axis=-1).tolist()
        bcs = [EquilibriumBC(rho=1.0, u=(0.02, 0.0, 0.0), indices=bne["top"]), FullwayBounceBackBC(indices=walls)]
        omega = 1.0
    elif case == "periodic":
        bcs, omega = [], 1.7
    else:  # flow past sphere (examples/cfd/flow_past_sphere_3d.py geometry)
        walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
        walls = np.unique(np.array(walls), axis=-1).tolist()
        X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        r = shape[1] // 6
        ind = np.where((X - shape[0] // 6) ** 2 + (Y - shape[1] // 2) ** 2 + (Z - shape[2] // 2) ** 2 < r**2)
        sph = [tuple(int(v) for v in ind[i]) for i in range(3)]
        Hy, Hz = float(shape[1] - 1), float(shape[2] - 1)

        def profile(index):
            yc, zc = index[1] - Hy / 2.0, index[2] - Hz / 2.0
            r2 = (2.0 * yc / Hy) ** 2.0 + (2.0 * zc / Hz) ** 2.0
            return [0.04 * np.maximum(0.0, 1.0 - r2)]

        bcs = [FullwayBounceBackBC(indices=walls), RegularizedBC("velocity", profile=profile, indices=bne["left"]),
               ExtrapolationOutflowBC(indices=bne["right"]), HalfwayBounceBackBC(indices=sph)]  # fmt: skip
        omega = 1.6
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=collision)
    return grid, stepper, omega


def run(case, shape, lattice, policy, collision, steps, distributed):
    grid, stepper, omega = build(case, shape, lattice, policy, collision, distributed)
    init = None
    if case == "periodic":
        def init(grid, velocity_set, precision_policy, compute_backend):
            from xlb_b200.helper import initialize_eq

            gen = torch.Generator().manual_seed(0)
            u_all = 0.02 * torch.randn((3,) + tuple(grid.shape), generator=gen)
            x0, nxl = grid.start_index[0], grid.local_shape[0]
            u = grid.create_field(3, dtype=precision_policy.compute_precision)
            u.copy_(u_all[:, x0 : x0 + nxl])
            rho = grid.create_field(1, dtype=precision_policy.compute_precision, fill_value=1.0)
            return initialize_eq(grid.create_field(velocity_set.q), grid, velocity_set, precision_policy, compute_backend, rho=rho, u=u)

    f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields(initializer=init)
    for i in range(steps):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, i)
        f_0, f_1 = f_1, f_0
    torch.cuda.synchronize()
    return f_0, bc_mask, missing_mask


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ok_all = True
    cases = [
        ("cavity", (16 * world, 24, 32), "D3Q19", "FP32FP32", "BGK", 25),
        ("periodic", (8 * world, 16, 32), "D3Q19", "FP32FP32", "BGK", 25),
        ("periodic", (8 * world, 16, 32), "D3Q27", "FP32FP16", "BGK", 15),
        ("sphere", (24 * world, 24, 24), "D3Q27", "FP32FP32", "KBC", 25),
        ("sphere", (24 * world, 24, 24), "D3Q19", "FP64FP32", "BGK", 25),
        ("cavity2d", (16 * world, 40), "D2Q9", "FP32FP32", "BGK", 40),
        ("cavity2d", (16 * world, 40), "D2Q9", "FP32FP32", "KBC", 40),
        ("cavity", (8 * world, 16, 64), "D3Q19", "FP32FP16", "BGK", 25),  # tile-kernel shape (nz | 512): interior planes take the tile path
        ("cavity", (8 * world, 16, 64), "D3Q19", "FP32FP32", "BGK", 25),  # ... the scalar tile kernel (fp32 storage)
        ("cavity", (8 * world, 16, 64), "D3Q19", "FP64FP32", "BGK", 25),
    ]
    for case, shape, lattice, policy, collision, steps in cases:
        f, bc, mm = run(case, shape, lattice, policy, collision, steps, None)
        parts = [torch.empty_like(f) for _ in range(world)]
        dist.all_gather(parts, f.contiguous())
        bparts = [torch.empty_like(bc) for _ in range(world)]
        dist.all_gather(bparts, bc.contiguous())
        mparts = [torch.empty_like(mm.to(torch.uint8)) for _ in range(world)]
        dist.all_gather(mparts, mm.to(torch.uint8).contiguous())
        if rank == 0:
            whole = torch.cat(parts, dim=1)
            ref, bc_ref, mm_ref = run(case, shape, lattice, policy, collision, steps, (0, 1))
            same = torch.equal(whole, ref)
            masks = torch.equal(torch.cat(bparts, dim=1), bc_ref) and torch.equal(torch.cat(mparts, dim=1).bool(), mm_ref)
            finite = bool(torch.isfinite(whole.float()).all())
            print(f"[mgpu x{world}] {case} {lattice} {collision} {policy} {shape} {steps} steps: bit-identical={same} masks={masks} finite={finite} "
                  f"maxdiff={float((whole.double() - ref.double()).abs().max()):.3e}", flush=True)  # fmt: skip
            ok_all &= same and masks and finite
        dist.barrier()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok_all else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
