"""Body-force driven turbulent channel in the style of the reference's examples/cfd/turbulent_channel_3d.py: periodic in x and
y, no-slip z walls as RegularizedBC("velocity", (0, 0, 0)), D3Q27 + KBC wrapped in ForcedCollision (ExactDifference), FP64FP64,
seeded random start through helper.initialize_eq(u=...).

    python examples/turbulent_channel.py [half_width] [steps]
"""

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
from xlb.helper import initialize_eq
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import RegularizedBC
from xlb.operator.macroscopic import Macroscopic

argv = [int(a) for a in sys.argv[1:]]
channel_half_width = argv[0] if argv else 50
num_steps = argv[1] if len(argv) > 1 else 2000
grid_shape = (6 * channel_half_width, 3 * channel_half_width, 2 * channel_half_width)
Re_tau, u_tau = 180, 0.001
visc = u_tau * channel_half_width / Re_tau
omega = 1.0 / (3.0 * visc + 0.5)
compute_backend, precision_policy = ComputeBackend.WARP, PrecisionPolicy.FP64FP64
velocity_set = xlb.velocity_set.D3Q27(precision_policy=precision_policy, compute_backend=compute_backend)
xlb.init(velocity_set=velocity_set, default_backend=compute_backend, default_precision_policy=precision_policy)
grid = grid_factory(grid_shape, compute_backend=compute_backend)

force_vector = np.zeros(velocity_set.d)
force_vector[0] = Re_tau**2 * visc**2 / channel_half_width**3
box = grid.bounding_box_indices(remove_edges=True)
walls = [box["bottom"][i] + box["top"][i] for i in range(velocity_set.d)]
boundary_conditions = [RegularizedBC("velocity", prescribed_value=(0.0, 0.0, 0.0), indices=walls)]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=boundary_conditions, collision_type="KBC", force_vector=force_vector)
f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()

np.random.seed(0)
u_init = wp.array(1e-2 * np.random.random((velocity_set.d,) + grid_shape), dtype=precision_policy.compute_precision.wp_dtype)
f_0 = initialize_eq(f_0, grid, velocity_set, precision_policy, compute_backend, u=u_init)
macro = Macroscopic(compute_backend=compute_backend, velocity_set=velocity_set, precision_policy=precision_policy)
rho = grid.create_field(cardinality=1, dtype=precision_policy.compute_precision)
u = grid.create_field(cardinality=velocity_set.d, dtype=precision_policy.compute_precision)

start = time.time()
for step in range(num_steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, step)
    f_0, f_1 = f_1, f_0
    if (step + 1) % max(1, num_steps // 5) == 0:
        rho, u = macro(f_0, rho, u)
        ux = u.numpy()[0]
        profile = ux.mean(axis=(0, 1))  # mean streamwise velocity over wall-normal position z
        print(f"step {step + 1}: bulk u_x {ux.mean():.5f}, centreline u_x {profile[grid_shape[2] // 2]:.5f}, u+ at centre {profile[grid_shape[2] // 2] / u_tau:.2f}")
elapsed = time.time() - start
print(f"{np.prod(grid_shape) * num_steps / elapsed / 1e6:.0f} MLUPS (D3Q27 KBC + force, FP64FP64)")
