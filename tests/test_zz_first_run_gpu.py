"""GPU cases that were added AFTER the round's GPU budget was spent.  They are pinned on the CPU (reference vectors vs
both oracles) but the CUDA path runs them here for the first time, so each one executes in its own subprocess (a device
fault cannot poison the test session) and is marked xfail(strict=False): a failure is information for the next round,
not a red suite.  Promote them into tests/test_native_step_gpu.py once they have been seen green."""

import os
import subprocess
import sys

import pytest

from common import EXTRA_CASES_2D

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
from common import load_golden, native_run, rel_err, unpack_bits
g = load_golden(%(name)r)
q = g["f_final"].shape[0]
f, bc_mask, missing = native_run(g, backend=%(backend)r)
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


@pytest.mark.xfail(strict=False, reason="first GPU execution of cases added after the round-1 GPU budget was spent")
@pytest.mark.parametrize("backend", ["WARP", "JAX"])
@pytest.mark.parametrize("name", EXTRA_CASES_2D)
def test_first_run_of_late_cases(name, backend):
    proc = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "name": name, "backend": backend}], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-1500:]
    lines = {l.split()[0]: l.split()[1:] for l in proc.stdout.splitlines() if l.startswith(("RESULT", "FORCE"))}
    assert lines["RESULT"][0] == "True", "masks differ"
    assert float(lines["RESULT"][1]) <= 1e-5, lines["RESULT"]
    if "FORCE" in lines:
        assert lines["FORCE"][0] == "True"
