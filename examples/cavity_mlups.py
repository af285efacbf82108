"""Lid-driven cavity throughput (MLUPS), written against the XLB operator API exactly as a user of Autodesk/XLB would
(same calls as the reference's examples/performance/mlups_3d.py: xlb.init, grid_factory, bounding_box_indices,
EquilibriumBC / FullwayBounceBackBC, IncompressibleNavierStokesStepper, prepare_fields, the step-and-swap loop,
wp.synchronize).  `import xlb` resolves to xlb_b200 through the alias package at the repository root.

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

parser = argparse.ArgumentParser()
parser.add_argument("cube_edge", type=int)
parser.add_argument("num_steps", type=int)
parser.add_argument("compute_backend", type=str, help="jax or warp: selects the call convention, both run the CUDA kernels")
parser.add_argument("precision", type=str, help="fp32/fp32, fp64/fp64, fp64/fp32, fp32/fp16")
args = parser.parse_args()

backend = ComputeBackend.JAX if args.compute_backend == "jax" else ComputeBackend.WARP
policy = {"fp32/fp32": PrecisionPolicy.FP32FP32, "fp64/fp64": PrecisionPolicy.FP64FP64, "fp64/fp32": PrecisionPolicy.FP64FP32, "fp32/fp16": PrecisionPolicy.FP32FP16}[args.precision]
xlb.init(velocity_set=xlb.velocity_set.D3Q19(precision_policy=policy, compute_backend=backend), default_backend=backend, default_precision_policy=policy)

n = args.cube_edge
grid = grid_factory((n, n, n))
box = grid.bounding_box_indices()
box_no_edge = grid.bounding_box_indices(remove_edges=True)
lid = box_no_edge["top"]
walls = [box["bottom"][i] + box["left"][i] + box["right"][i] + box["front"][i] + box["back"][i] for i in range(3)]
walls = np.unique(np.array(walls), axis=-1).tolist()
boundary_conditions = [EquilibriumBC(rho=1.0, u=(0.02, 0.0, 0.0), indices=lid), FullwayBounceBackBC(indices=walls)]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=boundary_conditions, collision_type="BGK")
if backend == ComputeBackend.JAX:
    stepper = distribute(stepper, grid, xlb.velocity_set.D3Q19(precision_policy=policy, compute_backend=backend))

omega = 1.0
f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()
for i in range(5):  # warm-up (the reference script times its JIT compilation too; nothing is compiled here)
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, i)
    f_0, f_1 = f_1, f_0
wp.synchronize()
start = time.time()
for i in range(args.num_steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, i)
    f_0, f_1 = f_1, f_0
wp.synchronize()
elapsed = time.time() - start
print(f"Simulation completed in {elapsed:.2f} seconds")
print(f"MLUPs: {n**3 * args.num_steps / elapsed / 1e6:.2f}")
