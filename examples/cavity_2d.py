"""2-D lid-driven cavity (D2Q9) in the style of the reference's examples/cfd/lid_driven_cavity_2d.py: EquilibriumBC lid,
HalfwayBounceBack walls, stepper on the WARP convention, post-processing through a JAX-convention Macroscopic on
`wp.to_jax(f_0)`.

    python examples/cavity_2d.py [n] [steps] [BGK|KBC]
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import xlb
import warp as wp
import jax.numpy as jnp
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import HalfwayBounceBackBC, EquilibriumBC
from xlb.operator.macroscopic import Macroscopic
from xlb.utils import save_image

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
num_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
collision = sys.argv[3] if len(sys.argv) > 3 else "BGK"
compute_backend, precision_policy = ComputeBackend.WARP, PrecisionPolicy.FP32FP32
velocity_set = xlb.velocity_set.D2Q9(precision_policy=precision_policy, compute_backend=compute_backend)
xlb.init(velocity_set=velocity_set, default_backend=compute_backend, default_precision_policy=precision_policy)
grid = grid_factory((n, n), compute_backend=compute_backend)

box = grid.bounding_box_indices()
box_no_edge = grid.bounding_box_indices(remove_edges=True)
lid = box_no_edge["top"]
walls = [box["bottom"][i] + box["left"][i] + box["right"][i] for i in range(velocity_set.d)]
walls = np.unique(np.array(walls), axis=-1).tolist()

u_lid, Re = 0.05, 200.0
omega = 1.0 / (3.0 * (u_lid * (n - 1) / Re) + 0.5)
boundary_conditions = [EquilibriumBC(rho=1.0, u=(u_lid, 0.0), indices=lid), HalfwayBounceBackBC(indices=walls)]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=boundary_conditions, collision_type=collision)
f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()
for step in range(num_steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, step)
    f_0, f_1 = f_1, f_0
wp.synchronize()

macro = Macroscopic(
    compute_backend=ComputeBackend.JAX,
    precision_policy=precision_policy,
    velocity_set=xlb.velocity_set.D2Q9(precision_policy=precision_policy, compute_backend=ComputeBackend.JAX),
)
f_current = wp.to_jax(f_0)[..., 0]  # drop the trailing singleton of the WARP 2-D layout
rho, u = macro(f_current)
u_magnitude = jnp.sqrt(u[0] ** 2 + u[1] ** 2)
name = save_image(u_magnitude[1:-1, 1:-1], timestep=num_steps, prefix=os.environ.get("XLB_OUT_PREFIX", "/tmp/cavity2d_umag"))
print(f"D2Q9 {collision} {n}x{n}, {num_steps} steps, omega = {omega:.4f}: max |u| = {float(u_magnitude.max()):.4f}, mean rho = {float(rho.mean()):.6f}; wrote {name}")
assert bool(jnp.isnan(u_magnitude).sum() == 0)
