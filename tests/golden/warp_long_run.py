"""One-off check, not part of the test suite (7 minutes): the reference's WARP backend, interpreted per cell, on the mlups_3d.py
cavity (C1 of BASELINE.json in miniature: D3Q19 BGK, 8^3, 1000 steps), against the C oracle.  Run in the build container:

    cd /tmp && python /root/repo/tests/golden/warp_long_run.py

Recorded result (round 1): "C oracle bit-identical = True, max abs diff 0.000e+00" after 1000 steps."""

import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import refshim
xlb = refshim.import_reference("/root/reference", interpret_warp=True)
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, EquilibriumBC
from oracle import lbm_numpy as O, lbm_c
pp, be = PrecisionPolicy.FP32FP32, ComputeBackend.WARP
xlb.init(velocity_set=xlb.velocity_set.D3Q19(precision_policy=pp, compute_backend=be), default_backend=be, default_precision_policy=pp)
n, steps = 8, 1000
grid = grid_factory((n, n, n))
box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
walls = [box["bottom"][i] + box["left"][i] + box["right"][i] + box["front"][i] + box["back"][i] for i in range(3)]
walls = np.unique(np.array(walls), axis=-1).tolist()
bcs = [EquilibriumBC(rho=1.0, u=(0.02, 0.0, 0.0), indices=bne["top"]), FullwayBounceBackBC(indices=walls)]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="BGK")
f_0, f_1, bc_mask, missing = stepper.prepare_fields()
t = time.time()
for i in range(steps):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, 1.0, i)
    f_0, f_1 = f_1, f_0
lat = O.Lattice("D3Q19")
obcs = [O.BC("equilibrium", bcs[0].id, np.array(bne["top"]), rho=1.0, u=(0.02, 0, 0)), O.BC("fullway", bcs[1].id, np.array(walls))]
bm, mm = O.build_masks(obcs, (n, n, n), lat, flavor="warp")
f = lbm_c.run(O.initialize_eq((n, n, n), lat), bm, mm, obcs, 1.0, lat, steps, "FP32FP32", "BGK")
ref = np.asarray(f_0)
print("C1 in miniature: mlups_3d.py cavity D3Q19 BGK %d^3 x %d steps on the reference's WARP backend (interpreted) in %.0f s: C oracle bit-identical = %s, max abs diff %.3e" % (n, steps, time.time() - t, np.array_equal(f, ref), np.abs(f - ref).max()))
