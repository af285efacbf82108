"""The C/OpenMP oracle (oracle/lbm_ref.c, a per-cell restatement of the reference's fused Warp kernel) against the vectors
produced by the reference's own Python and against the numpy oracle.  CPU only."""

import numpy as np
import pytest

from common import LATE_CASES, STEP_CASES, load_golden, oracle_bcs, rel_err
from oracle import lbm_c
from oracle import lbm_numpy as O

pytestmark = pytest.mark.skipif(not lbm_c.available(), reason="oracle/liblbm_ref.so not built (make -C oracle)")


@pytest.mark.parametrize("name", STEP_CASES + LATE_CASES)
def test_c_oracle_matches_reference_vectors(name):
    g = load_golden(name)
    lat = O.Lattice(g["lattice"])
    bcs = oracle_bcs(g)
    if bcs:
        bc_mask, missing = O.build_masks(bcs, g["shape"], lat, flavor="warp")
    else:
        bc_mask, missing = np.zeros((1,) + g["shape"], np.uint8), np.zeros((lat.q,) + g["shape"], bool)
    f = lbm_c.run(g["f_init"], bc_mask, missing, bcs, g["omega"], lat, g["steps"], g["policy"], g["collision"])
    assert f.dtype == g["f_final"].dtype
    tol = {"FP32FP32": 2e-6, "FP64FP32": 2e-6, "FP64FP64": 1e-12, "FP32FP16": 1e-3}[g["policy"]]
    assert rel_err(f, g["f_final"]) <= tol


def test_c_oracle_equals_numpy_oracle_on_a_fresh_case():
    lat, shape, bcs, bc_mask, missing = lbm_c.cavity_case("D3Q19", 20, "FP32FP32")
    f0 = O.initialize_eq(shape, lat)
    a = lbm_c.run(f0, bc_mask, missing, bcs, 1.2, lat, 40)
    b = O.run(f0, bc_mask, missing, bcs, 1.2, lat, 40)
    assert rel_err(a, b) <= 1e-6
