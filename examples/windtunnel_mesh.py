"""Wind tunnel with a body given as a triangle mesh, in the style of the reference's examples/cfd/windtunnel_3d.py: Fullway
walls, RegularizedBC velocity inlet, ExtrapolationOutflowBC outlet, HalfwayBounceBackBC on the mesh (voxelised by
MeshBoundaryMasker), D3Q27 + KBC, drag from MomentumTransfer.

    python examples/windtunnel_mesh.py [body.stl] [nx ny nz] [steps]

Without an STL file an ellipsoid is triangulated on the fly.  The mesh is scaled to a quarter of the tunnel length and
placed a quarter of the way in, on the floor's mid-line, as windtunnel_3d.py:80-89 does.
"""

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import xlb
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC, ExtrapolationOutflowBC
from xlb.operator.force.momentum_transfer import MomentumTransfer
from xlb.utils import read_stl


def ellipsoid(n_lat=24, n_lon=48, radii=(1.0, 0.45, 0.35)):
    """Closed triangle soup (3 T, 3) of an ellipsoid."""
    th = np.linspace(0.0, np.pi, n_lat + 1)
    ph = np.linspace(0.0, 2.0 * np.pi, n_lon + 1)
    P = lambda i, j: np.array([radii[0] * np.cos(th[i]), radii[1] * np.sin(th[i]) * np.cos(ph[j]), radii[2] * np.sin(th[i]) * np.sin(ph[j])])
    tris = []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b, c, d = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
            if i > 0:
                tris += [a, b, d]
            if i < n_lat - 1:
                tris += [b, c, d]
    return np.array(tris)


args = sys.argv[1:]
stl = args.pop(0) if args and not args[0].isdigit() else None
nums = [int(a) for a in args]
grid_shape = tuple(nums[:3]) if len(nums) >= 3 else (256, 96, 96)
num_steps = nums[3] if len(nums) >= 4 else 2000
wind_speed, omega = 0.02, 1.9
compute_backend, precision_policy = ComputeBackend.WARP, PrecisionPolicy.FP32FP32
velocity_set = xlb.velocity_set.D3Q27(precision_policy=precision_policy, compute_backend=compute_backend)
xlb.init(velocity_set=velocity_set, default_backend=compute_backend, default_precision_policy=precision_policy)
grid = grid_factory(grid_shape, compute_backend=compute_backend)

box = grid.bounding_box_indices()
box_no_edge = grid.bounding_box_indices(remove_edges=True)
inlet, outlet = box_no_edge["left"], box_no_edge["right"]
walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(velocity_set.d)]
walls = np.unique(np.array(walls), axis=-1).tolist()

mesh_vertices = read_stl(stl) if stl else ellipsoid()
mesh_vertices = mesh_vertices - mesh_vertices.min(axis=0)
extents = mesh_vertices.max(axis=0)
dx = extents.max() / (grid_shape[0] / 4)
mesh_vertices = mesh_vertices / dx
body = mesh_vertices + np.array([grid_shape[0] / 4, (grid_shape[1] - extents[1] / dx) / 2, 2.0])
cross_section = np.prod(extents[1:]) / dx**2

bc_body = HalfwayBounceBackBC(mesh_vertices=body)
boundary_conditions = [
    FullwayBounceBackBC(indices=walls),
    RegularizedBC("velocity", prescribed_value=(wind_speed, 0.0, 0.0), indices=inlet),
    ExtrapolationOutflowBC(indices=outlet),
    bc_body,
]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=boundary_conditions, collision_type="KBC")
f_0, f_1, bc_mask, missing_mask = stepper.prepare_fields()
print(f"{body.shape[0] // 3} triangles -> {int((bc_mask.numpy() == 255).sum())} solid voxels, {int((bc_mask.numpy() == bc_body.id).sum())} boundary cells")
momentum_transfer = MomentumTransfer(bc_body, compute_backend=compute_backend)

start = time.time()
for step in range(num_steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, step)
    f_0, f_1 = f_1, f_0
    if (step + 1) % max(1, num_steps // 10) == 0:
        force = np.asarray(momentum_transfer(f_0, f_1, bc_mask, missing_mask))
        cd = 2.0 * force[0] / (wind_speed**2 * cross_section)
        print(f"step {step + 1}: drag coefficient {cd:.4f}, lift coefficient {2.0 * force[2] / (wind_speed**2 * cross_section):.4f}")
elapsed = time.time() - start
print(f"{np.prod(grid_shape) * num_steps / elapsed / 1e6:.0f} MLUPS incl. the force evaluations")
