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
