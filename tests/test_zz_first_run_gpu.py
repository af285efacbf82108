"""GPU cases that were added AFTER the round's GPU budget was spent.  They are pinned on the CPU (reference vectors vs
both oracles) but the CUDA path runs them here for the first time, so each one executes in its own subprocess (a device
fault cannot poison the test session) and is marked xfail(strict=False): a failure is information for the next round,
not a red suite.  Promote them into tests/test_native_step_gpu.py once they have been seen green."""

import os
import subprocess
import sys

import pytest

from common import EXTRA_CASES_2D, STEP_CASES, WARP_CASES, WARP_CASES_N4

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
from common import load_golden, native_run, rel_err, unpack_bits
g = load_golden(%(name)r)
q = g["f_final"].shape[0]
f, bc_mask, missing = native_run(g, backend=%(backend)r, cells_per_thread=%(v)d)
ok_masks = np.array_equal(bc_mask.reshape(g["bc_mask"].shape), g["bc_mask"]) and np.array_equal(missing.reshape((q,) + g["shape"]), unpack_bits(g["missing_bits"], q))
print("RESULT", ok_masks, rel_err(f, g["f_final"]))
if "force" in g:
    import torch
    from common import native_case
    from xlb_b200.operator.force import MomentumTransfer
    stepper, f_0, f_1, bm, mm = native_case(g, backend=%(backend)r)
    f_0.copy_(torch.as_tensor(g["f_final"]).reshape(f_0.shape))
    force = MomentumTransfer(stepper.boundary_conditions[int(g["force_bc"])])(f_0, f_1, bm, mm)
    force = np.asarray(force.numpy() if hasattr(force, "numpy") and not isinstance(force, np.ndarray) else force)
    print("FORCE", bool(np.allclose(force, g["force"], rtol=2e-5, atol=1e-7)))
"""


def run_child(name, backend, v=0):
    proc = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "name": name, "backend": backend, "v": v}], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-1500:]
    lines = {l.split()[0]: l.split()[1:] for l in proc.stdout.splitlines() if l.startswith(("RESULT", "FORCE"))}
    assert lines["RESULT"][0] == "True", "masks differ"
    assert float(lines["RESULT"][1]) <= 1e-5, lines["RESULT"]
    if "FORCE" in lines:
        assert lines["FORCE"][0] == "True"


LATE = pytest.mark.xfail(strict=False, reason="first GPU execution of cases added after the round-1 GPU budget was spent")


@LATE
@pytest.mark.parametrize("backend", ["WARP", "JAX"])
@pytest.mark.parametrize("name", EXTRA_CASES_2D)
def test_first_run_of_late_cases(name, backend):
    run_child(name, backend)


@LATE
@pytest.mark.parametrize("name", WARP_CASES)
def test_first_run_against_the_reference_warp_backend(name):
    """Vectors from the reference's own WARP backend (tests/golden/make_golden_warp.py); validated kernels, new fixtures."""
    run_child(name, "WARP")


@LATE
@pytest.mark.parametrize("name", WARP_CASES_N4)
def test_first_run_of_the_extended_collision_kernels(name):
    """SmagorinskyLESBGK / ForcedCollision in the fused step: kernels that have never run on a GPU."""
    run_child(name, "WARP")


@LATE
@pytest.mark.parametrize("name", [n for n in STEP_CASES + WARP_CASES if "kbc" in n])
def test_first_run_of_the_lean_kbc_variant(name):
    """cells_per_thread = 301 (register-lean KBC): host-validated, never run on a GPU."""
    run_child(name, "WARP", v=301)


OPS_CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch
import xlb_b200 as xlb
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision import BGK, KBC, ForcedCollision, SmagorinskyLESBGK
from xlb_b200.operator.force import ExactDifference
from oracle import lbm_numpy as O
from common import rel_err
pp, be = xlb.PrecisionPolicy.FP32FP32, ComputeBackend.WARP
worst = 0.0
for lattice in ("D3Q19", "D3Q27", "D2Q9"):
    vs = getattr(xlb.velocity_set, lattice)(pp, be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    lat = O.Lattice(lattice)
    shape = (6, 5, 4) if lat.d == 3 else (9, 7)
    rng = np.random.default_rng(3)
    rho = (1.0 + 0.01 * rng.standard_normal((1,) + shape)).astype(np.float32)
    u = (0.03 * rng.standard_normal((lat.d,) + shape)).astype(np.float32)
    feq = O.equilibrium(rho, u, lat)
    f = (feq * (1.0 + 0.02 * rng.standard_normal(feq.shape))).astype(np.float32)
    force = np.array([2e-4, -1e-4, 5e-5][: lat.d])
    dev = lambda a: torch.as_tensor(a if lat.d == 3 else a[..., None]).cuda()
    back = lambda t: t.cpu().numpy() if lat.d == 3 else t.cpu().numpy()[..., 0]
    F, FEQ, RHO, U = dev(f), dev(feq), dev(rho), dev(u)
    want = O.exact_difference_force(f.copy(), feq, rho, u, force, lat)
    got = back(ExactDifference(force)(F, FEQ, torch.empty_like(F), RHO, U))
    worst = max(worst, rel_err(got, want))
    cases = [("BGK", BGK)] + ([("KBC", KBC)] if lattice != "D3Q19" else []) + ([("SmagorinskyLESBGK", SmagorinskyLESBGK)] if lat.d == 3 else [])
    for cname, cls in cases:
        if cname == "BGK":
            base = O.collide_bgk(f, feq, 1.7)
        elif cname == "KBC":
            base = O.collide_kbc(f, feq, rho, lat, 1.7)
        else:
            base = O.collide_smagorinsky(f, feq, lat, 1.7)
            got = back(cls()(F, FEQ, RHO, U, torch.empty_like(F), 1.7))
            worst = max(worst, rel_err(got, base))
        want = O.exact_difference_force(base, feq, rho, u, force, lat)
        got = back(ForcedCollision(cls(), force_vector=force)(F, FEQ, torch.empty_like(F), RHO, U, 1.7))
        worst = max(worst, rel_err(got, want))
print("RESULT", worst)
"""


@LATE
def test_first_run_of_the_extended_collision_operators():
    """xlbn_collide_ext / xlbn_exact_difference through the operator classes vs the numpy oracle on random states."""
    proc = subprocess.run([sys.executable, "-c", OPS_CHILD % {"root": ROOT}], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-1500:]
    worst = [float(l.split()[1]) for l in proc.stdout.splitlines() if l.startswith("RESULT")][0]
    assert worst <= 2e-6, worst


MESH_CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np, torch
import xlb_b200 as xlb
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.grid import grid_factory
from xlb_b200.operator.boundary_condition import ExtrapolationOutflowBC, FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC
from xlb_b200.operator.boundary_masker import MeshBoundaryMasker
from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper
from oracle import lbm_c
from oracle import lbm_numpy as O
from common import rel_err, unpack_bits
from test_mesh_masker import MESH_CASES, load_mesh_case
pp, be = xlb.PrecisionPolicy.FP32FP32, ComputeBackend.WARP
ok = True
for name in MESH_CASES:
    g = load_mesh_case(name)
    lattice, shape = str(g["lattice"]), tuple(int(s) for s in g["shape"])
    vs = getattr(xlb.velocity_set, lattice)(pp, be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    grid = grid_factory(shape)
    lat = O.Lattice(lattice)
    for mode in ("reference", "schwarz_seidel"):
        bc = HalfwayBounceBackBC(mesh_vertices=g["vertices"].copy())
        bc.id = int(g["bc_id"])
        bc_mask = grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
        missing = grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
        bc_mask, missing = MeshBoundaryMasker(vs, pp, be, edge_test=mode)(bc, bc_mask, missing)
        bm, mm = O.build_masks_mesh(g["vertices"], bc.id, np.zeros((1,) + shape, np.uint8), np.zeros((lat.q,) + shape, bool), lat, edge_test=mode)
        same = np.array_equal(bc_mask.numpy(), bm) and np.array_equal(missing.numpy(), mm)
        if mode == "reference":
            same = same and np.array_equal(bc_mask.numpy(), g["bc_mask"]) and np.array_equal(missing.numpy(), unpack_bits(g["missing_bits"], lat.q))
        ok = ok and same
        print("MASK", name, mode, same)
# a wind-tunnel style run with a mesh body (examples/cfd/windtunnel_3d.py:66-96), against the C oracle on the same masks
g = load_mesh_case("warp_mesh_octahedron_d3q27")
shape = (24, 11, 10)
vs = xlb.velocity_set.D3Q27(pp, be)
xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
grid = grid_factory(shape)
box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
walls = np.unique(np.array(walls), axis=-1).tolist()
bcs = [FullwayBounceBackBC(indices=walls), RegularizedBC("velocity", prescribed_value=(0.03, 0.0, 0.0), indices=bne["left"]),
       ExtrapolationOutflowBC(indices=bne["right"]), HalfwayBounceBackBC(mesh_vertices=g["vertices"] + np.array([2.0, 0.0, 0.0]))]
stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="KBC")
f_0, f_1, bc_mask, missing = stepper.prepare_fields()
lat = O.Lattice("D3Q27")
obcs = [O.BC("fullway", bcs[0].id, np.array(walls)), O.BC("regularized", bcs[1].id, np.array(bne["left"]), bc_type="velocity", prescribed=np.array([0.03, 0.0, 0.0])),
        O.BC("outflow", bcs[2].id, np.array(bne["right"])), O.BC("halfway", bcs[3].id, np.zeros((3, 0), np.int64))]
bm, mm = O.build_masks(obcs[:3], shape, lat, flavor="warp")
bm, mm = O.build_masks_mesh(g["vertices"] + np.array([2.0, 0.0, 0.0]), bcs[3].id, bm, mm, lat)
same = np.array_equal(bc_mask.numpy(), bm) and np.array_equal(missing.numpy(), mm)
print("MASK windtunnel", same)
for i in range(20):
    f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, 1.6, i)
    f_0, f_1 = f_1, f_0
ref = lbm_c.run(O.initialize_eq(shape, lat), bm, mm, obcs, 1.6, lat, 20, "FP32FP32", "KBC")
err = rel_err(f_0.numpy(), ref)
print("RESULT", ok and same, err)
"""


@LATE
def test_first_run_of_the_mesh_boundary_masker():
    """xlbn_mask_mesh through MeshBoundaryMasker: both edge tests vs the oracle (the literal one also vs the reference's masks),
    and a wind-tunnel run with a mesh body vs the C oracle on the same masks."""
    proc = subprocess.run([sys.executable, "-c", MESH_CHILD % {"root": ROOT}], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-1500:]
    res = [l.split() for l in proc.stdout.splitlines() if l.startswith("RESULT")][0]
    assert res[1] == "True", proc.stdout[-1500:]
    assert float(res[2]) <= 1e-5
