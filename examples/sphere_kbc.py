"""Flow past a sphere, D3Q27 + KBC, in the style of the reference's examples/cfd/flow_past_sphere_3d.py: Fullway walls,
RegularizedBC velocity inlet with a Poiseuille profile given as a per-cell `@wp.func`, ExtrapolationOutflowBC outlet,
HalfwayBounceBackBC sphere, stepper on the WARP convention and post-processing with a JAX-convention Macroscopic.

    python examples/sphere_kbc.py [nx ny nz] [steps]
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import xlb
import warp as wp
import jax.numpy as jnp
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC, ExtrapolationOutflowBC
from xlb.operator.macroscopic import Macroscopic
from xlb.utils import save_image

argv = [int(a) for a in sys.argv[1:]]
grid_shape = tuple(argv[:3]) if len(argv) >= 3 else (256, 64, 64)
num_steps = argv[3] if len(argv) >= 4 else 2000
omega, u_max = 1.6, 0.04
compute_backend, precision_policy = ComputeBackend.WARP, PrecisionPolicy.FP32FP32
velocity_set = xlb.velocity_set.D3Q27(precision_policy=precision_policy, compute_backend=compute_backend)
xlb.init(velocity_set=velocity_set, default_backend=compute_backend, default_precision_policy=precision_policy)
grid = grid_factory(grid_shape, compute_backend=compute_backend)

box = grid.bounding_box_indices()
box_no_edge = grid.bounding_box_indices(remove_edges=True)
inlet, outlet = box_no_edge["left"], box_no_edge["right"]
walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(velocity_set.d)]
walls = np.unique(np.array(walls), axis=-1).tolist()
radius = grid_shape[1] // 12
X, Y, Z = np.meshgrid(*[np.arange(s) for s in grid_shape], indexing="ij")
inside = np.where((X - grid_shape[0] // 6) ** 2 + (Y - grid_shape[1] // 2) ** 2 + (Z - grid_shape[2] // 2) ** 2 < radius**2)
sphere = [tuple(inside[i]) for i in range(velocity_set.d)]
H_y, H_z = float(grid_shape[1] - 1), float(grid_shape[2] - 1)


@wp.func
def inlet_profile(index: wp.vec3i):
    y, z = wp.float32(index[1]), wp.float32(index[2])
    r_squared = (2.0 * (y - H_y / 2.0) / H_y) ** 2.0 + (2.0 * (z - H_z / 2.0) / H_z) ** 2.0
    return wp.vec(u_max * wp.max(0.0, 1.0 - r_squared), length=1)


boundary_conditions = [
    FullwayBounceBackBC(indices=walls),
    RegularizedBC("velocity", profile=inlet_profile, indices=inlet),
    ExtrapolationOutflowBC(indices=outlet),
    HalfwayBounceBackBC(indices=sphere),
]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=boundary_conditions, collision_type="KBC")
f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()
macro = Macroscopic(
    compute_backend=ComputeBackend.JAX,
    precision_policy=precision_policy,
    velocity_set=xlb.velocity_set.D3Q27(precision_policy=precision_policy, compute_backend=ComputeBackend.JAX),
)

start = time.time()
for step in range(num_steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, step)
    f_0, f_1 = f_1, f_0
wp.synchronize()
elapsed = time.time() - start
f_current = f_0 if isinstance(f_0, jnp.ndarray) else wp.to_jax(f_0)
rho, u = macro(f_current)
u = u[:, 1:-1, 1:-1, 1:-1]
u_magnitude = jnp.sqrt(u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
name = save_image(u_magnitude[:, grid_shape[1] // 2, :], timestep=num_steps, prefix=os.environ.get("XLB_OUT_PREFIX", "/tmp/sphere_umag"))
cells = float(np.prod(grid_shape))
print(f"{num_steps} steps in {elapsed:.2f} s -> {cells * num_steps / elapsed / 1e6:.1f} MLUPS; max |u| = {float(u_magnitude.max()):.4f}; wrote {name}")
assert bool(jnp.isnan(u_magnitude).sum() == 0), "NaN in the velocity field"
