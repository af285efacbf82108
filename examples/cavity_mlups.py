"""Lid-driven cavity throughput (MLUPS), written against the XLB operator API as a user of Autodesk/XLB would (the workflow of
the reference's examples/performance/mlups_3d.py: xlb.init, grid_factory, bounding_box_indices, EquilibriumBC /
FullwayBounceBackBC, IncompressibleNavierStokesStepper, prepare_fields, the step-and-swap loop, wp.synchronize).
`import xlb` resolves to xlb_b200 through the alias package at the repository root.

    python examples/cavity_mlups.py 512 200 warp fp32/fp32
"""

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import xlb
import warp as wp
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, EquilibriumBC
from xlb.distribute import distribute

POLICIES = {"fp32/fp32": PrecisionPolicy.FP32FP32, "fp64/fp64": PrecisionPolicy.FP64FP64, "fp64/fp32": PrecisionPolicy.FP64FP32, "fp32/fp16": PrecisionPolicy.FP32FP16}
LID_VELOCITY, OMEGA, WARMUP_STEPS = (0.02, 0.0, 0.0), 1.0, 5


def command_line():
    cli = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    cli.add_argument("cube_edge", type=int)
    cli.add_argument("num_steps", type=int)
    cli.add_argument("compute_backend", choices=["jax", "warp"], help="selects the call convention; both run the CUDA kernels")
    cli.add_argument("precision", choices=sorted(POLICIES))
    return cli.parse_args()


def cavity(edge, backend, policy):
    """Closed box: moving lid (top face without its rim) as EquilibriumBC, the other five faces as full-way bounce-back."""
    lattice = xlb.velocity_set.D3Q19(precision_policy=policy, compute_backend=backend)
    xlb.init(velocity_set=lattice, default_backend=backend, default_precision_policy=policy)
    grid = grid_factory((edge,) * 3)
    faces, faces_without_rim = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    wall_cells = [sum((faces[side][axis] for side in ("bottom", "left", "right", "front", "back")), []) for axis in range(3)]
    wall_cells = np.unique(np.array(wall_cells), axis=-1).tolist()
    bcs = [EquilibriumBC(rho=1.0, u=LID_VELOCITY, indices=faces_without_rim["top"]), FullwayBounceBackBC(indices=wall_cells)]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="BGK")
    if backend == ComputeBackend.JAX:  # the reference shards its JAX stepper explicitly; here the grid already knows its slab
        stepper = distribute(stepper, grid, lattice)
    return stepper


def advance(stepper, fields, steps, first_step=0):
    f_now, f_next, bc_mask, missing_mask = fields
    for t in range(first_step, first_step + steps):
        f_now, f_next = stepper(f_now, f_next, bc_mask, missing_mask, OMEGA, t)
        f_now, f_next = f_next, f_now
    wp.synchronize()
    return f_now, f_next, bc_mask, missing_mask


if __name__ == "__main__":
    args = command_line()
    backend = ComputeBackend.JAX if args.compute_backend == "jax" else ComputeBackend.WARP
    stepper = cavity(args.cube_edge, backend, POLICIES[args.precision])
    fields = advance(stepper, stepper.prepare_fields(), WARMUP_STEPS)  # the reference script times its JIT compilation too; nothing compiles here
    start = time.time()
    advance(stepper, fields, args.num_steps, first_step=WARMUP_STEPS)
    elapsed = time.time() - start
    print(f"Simulation completed in {elapsed:.2f} seconds")
    print(f"MLUPs: {args.cube_edge**3 * args.num_steps / elapsed / 1e6:.2f}")
