"""The two CPU oracles are independent restatements of the reference (whole-array numpy after the JAX path, per-cell C
after the Warp kernel).  Seeded random set-ups — random extents, relaxation rates, obstacles and boundary sets — must give
the same populations from both; and, where /root/reference is mounted (build container only), the reference's own
Python run live — its JAX backend under the numpy `jax` stand-in, its WARP backend under the interpretive `warp` stand-in —
must agree with them too."""

import os
import subprocess
import sys

import numpy as np
import pytest

from common import rel_err
from oracle import lbm_c
from oracle import lbm_numpy as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_c = pytest.mark.skipif(not lbm_c.available(), reason="oracle/liblbm_ref.so not built (make -C oracle)")


def random_tunnel(seed, lattice):
    rng = np.random.default_rng(seed)
    lat = O.Lattice(lattice)
    shape = (int(rng.integers(18, 30)), int(rng.integers(10, 16)), int(rng.integers(10, 16)))
    box, bne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
    walls = np.unique(np.concatenate([box[k] for k in ("bottom", "top", "front", "back")], axis=1), axis=-1)
    solid = np.zeros(shape, dtype=bool)
    for _ in range(int(rng.integers(1, 4))):  # a few boxes strictly inside the tunnel
        lo = [int(rng.integers(3, s - 6)) for s in shape]
        hi = [l + int(rng.integers(1, 4)) for l in lo]
        solid[lo[0] : hi[0], lo[1] : hi[1], lo[2] : hi[2]] = True
    body = np.array(np.nonzero(solid))
    inlet_kind = ["regularized", "zouhe"][int(rng.integers(0, 2))]
    outlet = [("outflow", {}), ("zouhe", dict(bc_type="pressure", prescribed=np.float64(1.0))), ("donothing", {})][int(rng.integers(0, 3))]
    u_in = float(rng.uniform(0.01, 0.04))
    bcs = [
        O.BC("fullway", 1, walls),
        O.BC(inlet_kind, 2, bne["left"], bc_type="velocity", prescribed=np.array([u_in, 0.0, 0.0])),
        O.BC(outlet[0], 3, bne["right"], **outlet[1]),
        O.BC("halfway", 4, body),
    ]
    omega = float(rng.uniform(1.0, 1.7))
    return lat, shape, bcs, omega


@needs_c
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("lattice,collision", [("D3Q19", "BGK"), ("D3Q27", "KBC")])
def test_numpy_and_c_oracles_agree_on_random_tunnels(seed, lattice, collision):
    lat, shape, bcs, omega = random_tunnel(seed, lattice)
    bm_w, mm_w = O.build_masks(bcs, shape, lat, flavor="warp")
    bm_j, mm_j = O.build_masks(bcs, shape, lat, flavor="jax")
    assert np.array_equal(bm_w, bm_j) and np.array_equal(mm_w, mm_j), "the two masker algorithms agree on bounded domains"
    f0 = O.initialize_eq(shape, lat)
    a = O.run(f0, bm_j, mm_j, bcs, omega, lat, 25, collision=collision)
    b = lbm_c.run(f0, bm_w, mm_w, bcs, omega, lat, 25, "FP32FP32", collision)
    assert np.isfinite(a).all() and rel_err(b, a) <= 5e-6, rel_err(b, a)


LIVE = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from oracle import refshim
xlb = refshim.import_reference("/root/reference")
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, EquilibriumBC
from oracle import lbm_numpy as O
pp, be = PrecisionPolicy.FP32FP32, ComputeBackend.JAX
xlb.init(velocity_set=xlb.velocity_set.D3Q27(precision_policy=pp, compute_backend=be), default_backend=be, default_precision_policy=pp)
n = 12
grid = grid_factory((n, n, n))
box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
walls = [box["bottom"][i] + box["left"][i] + box["right"][i] + box["front"][i] + box["back"][i] for i in range(3)]
walls = np.unique(np.array(walls), axis=-1).tolist()
bcs = [EquilibriumBC(rho=1.0, u=(0.03, 0.0, 0.0), indices=bne["top"]), FullwayBounceBackBC(indices=walls)]
ids = [b.id for b in bcs]
lid, wl = np.array(bne["top"]), np.array(walls)
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="KBC")
f_0, f_1, bc_mask, missing = stepper.prepare_fields()
for i in range(15):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, 1.7, i)
    f_0, f_1 = f_1, f_0
lat = O.Lattice("D3Q27")
obcs = [O.BC("equilibrium", ids[0], lid, rho=1.0, u=(0.03, 0, 0)), O.BC("fullway", ids[1], wl)]
bm, mm = O.build_masks(obcs, (n, n, n), lat, flavor="jax")
f = O.run(O.initialize_eq((n, n, n), lat), bm, mm, obcs, 1.7, lat, 15, collision="KBC")
ref = np.asarray(f_0)
print("MASKS", np.array_equal(bm, np.asarray(bc_mask)) and np.array_equal(mm, np.asarray(missing)))
print("RELERR", float(np.abs(f - ref).max() / np.abs(ref).max()))
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/xlb"), reason="the reference is only mounted in the build container")
def test_reference_run_live_matches_the_oracle():
    """Not a fixture: executes /root/reference's stepper (JAX backend under oracle/refshim) right now, in a subprocess."""
    proc = subprocess.run([sys.executable, "-c", LIVE % {"root": ROOT}], capture_output=True, text=True, timeout=300, cwd="/tmp")
    assert proc.returncode == 0, proc.stderr[-2000:]
    out = dict(line.split(" ", 1) for line in proc.stdout.splitlines() if line.startswith(("MASKS", "RELERR")))
    assert out["MASKS"].strip() == "True"
    assert float(out["RELERR"]) <= 1e-7


LIVE_WARP = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from oracle import refshim
xlb = refshim.import_reference("/root/reference", interpret_warp=True)
from xlb.compute_backend import ComputeBackend
from xlb.precision_policy import PrecisionPolicy
from xlb.grid import grid_factory
from xlb.operator.stepper import IncompressibleNavierStokesStepper
from xlb.operator.boundary_condition import FullwayBounceBackBC, ZouHeBC, ExtrapolationOutflowBC, HalfwayBounceBackBC
from oracle import lbm_numpy as O
from oracle import lbm_c
pp, be = PrecisionPolicy.FP32FP32, ComputeBackend.WARP
xlb.init(velocity_set=xlb.velocity_set.D3Q19(precision_policy=pp, compute_backend=be), default_backend=be, default_precision_policy=pp)
shape = (12, 7, 7)
grid = grid_factory(shape)
box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
walls = np.unique(np.array(walls), axis=-1).tolist()
block = [[4, 4, 5, 5], [3, 3, 3, 3], [3, 4, 3, 4]]
bcs = [FullwayBounceBackBC(indices=walls), ZouHeBC("velocity", prescribed_value=(0.03, 0.0, 0.0), indices=bne["left"]),
       ExtrapolationOutflowBC(indices=bne["right"]), HalfwayBounceBackBC(indices=block)]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="BGK")
f_0, f_1, bc_mask, missing = stepper.prepare_fields()
for i in range(6):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, 1.5, i)
    f_0, f_1 = f_1, f_0
lat = O.Lattice("D3Q19")
obcs = [O.BC("fullway", bcs[0].id, np.array(walls)), O.BC("zouhe", bcs[1].id, np.array(bne["left"]), bc_type="velocity", prescribed=np.array([0.03, 0.0, 0.0])),
        O.BC("outflow", bcs[2].id, np.array(bne["right"])), O.BC("halfway", bcs[3].id, np.array(block))]
bm, mm = O.build_masks(obcs, shape, lat, flavor="warp")
f = lbm_c.run(O.initialize_eq(shape, lat), bm, mm, obcs, 1.5, lat, 6, "FP32FP32", "BGK")
ref = np.asarray(f_0)
print("MASKS", np.array_equal(bm, np.asarray(bc_mask)) and np.array_equal(mm, np.asarray(missing)))
print("EXACT", np.array_equal(f, ref), float(np.abs(f - ref).max()))
"""


@needs_c
@pytest.mark.skipif(not os.path.isdir("/root/reference/xlb"), reason="the reference is only mounted in the build container")
def test_reference_warp_backend_live_equals_the_c_oracle():
    """Executes the reference's WARP backend right now (its fused kernel, BC functionals and Warp masker interpreted per
    cell by oracle/refshim's `warp` stand-in) and requires the C oracle — the restatement of that kernel — to match it
    bit for bit."""
    proc = subprocess.run([sys.executable, "-c", LIVE_WARP % {"root": ROOT}], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert proc.returncode == 0, proc.stderr[-2000:]
    out = dict(line.split(" ", 1) for line in proc.stdout.splitlines() if line.startswith(("MASKS", "EXACT")))
    assert out["MASKS"].strip() == "True"
    assert out["EXACT"].split()[0] == "True", out["EXACT"]
