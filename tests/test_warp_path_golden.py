"""Vectors produced by the reference's WARP backend — its fused step kernel, BC functionals, aux-data kernels, Warp
masker and MomentumTransfer executed cell by cell under oracle/refshim's interpretive `warp` stand-in
(tests/golden/make_golden_warp.py) — against the two oracles.  The C oracle restates exactly that kernel and reproduces it BIT
FOR BIT on every case (asserted); the numpy oracle follows the JAX path (different summation order) and is held to 2e-6."""

import numpy as np
import pytest

from common import RTOL, WARP_CASES, WARP_CASES_FP16, WARP_CASES_N4, c_oracle_run, load_golden, oracle_masks, oracle_run, rel_err, unpack_bits
from oracle import lbm_c
from oracle import lbm_numpy as O

needs_c = pytest.mark.skipif(not lbm_c.available(), reason="oracle/liblbm_ref.so not built (make -C oracle)")


def check_masks(g, bc_mask, missing):
    q = g["f_final"].shape[0]
    assert np.array_equal(bc_mask.reshape(g["bc_mask"].shape), g["bc_mask"])
    assert np.array_equal(missing, unpack_bits(g["missing_bits"], q))


@pytest.mark.parametrize("name", WARP_CASES + WARP_CASES_N4)
def test_numpy_oracle_matches_the_warp_backend(name):
    g = load_golden(name)
    assert g["backend"] == "WARP" and g["policy"] in ("FP32FP32", "FP64FP64")
    f, bc_mask, missing = oracle_run(g, flavor="warp")
    check_masks(g, bc_mask, missing)
    assert rel_err(f, g["f_final"]) <= (2e-6 if g["policy"] == "FP32FP32" else 1e-13)


@needs_c
@pytest.mark.parametrize("name", WARP_CASES_FP16)
def test_c_oracle_matches_the_warp_backend_under_fp16_storage(name):
    """FP32FP16 on the WARP backend: prescribed values rounded to fp16 in f_1[0, cell] (boundary_condition.py:151), every load widened,
    every store narrowed by the reference's explicit casts.  Bit for bit."""
    g = load_golden(name)
    assert g["backend"] == "WARP" and g["policy"] == "FP32FP16" and g["f_final"].dtype == np.float16
    f, bc_mask, missing = c_oracle_run(g)
    check_masks(g, bc_mask, missing)
    assert np.array_equal(f, g["f_final"]), rel_err(f, g["f_final"])


@needs_c
@pytest.mark.parametrize("name", WARP_CASES + WARP_CASES_N4)
def test_c_oracle_matches_the_warp_backend(name):
    g = load_golden(name)
    f, bc_mask, missing = c_oracle_run(g)
    check_masks(g, bc_mask, missing)
    assert np.array_equal(f, g["f_final"]), rel_err(f, g["f_final"])


@pytest.mark.parametrize("name", [n for n in WARP_CASES if "tunnel" in n])
def test_momentum_transfer_matches_the_warp_kernel(name):
    """MomentumTransfer.warp_implementation (momentum_transfer.py:92-176) on the final state of the 3-D tunnel cases."""
    g = load_golden(name)
    lat, bcs, bc_mask, missing = oracle_masks(g, "warp")
    force = O.momentum_transfer(bcs[int(g["force_bc"])], g["f_final"], bc_mask, missing, lat)
    # fp32 sums over ~100 edge cells in a different order: components are compared on the scale of the largest one
    assert np.allclose(force, g["force"], rtol=2e-5, atol=2e-6 * np.abs(g["force"]).max())


def test_solid_cells_are_never_touched():
    """bc_mask == 255 (nse_stepper.py:356-358): the Warp kernel returns before reading or writing anything."""
    g = load_golden("warp_cavity_d3q19_bgk_solid255")
    solid = tuple(g["solid255"])
    assert (g["bc_mask"][0][solid] == 255).all()
    assert np.array_equal(g["f_final"][(slice(None),) + solid], g["f_init"][(slice(None),) + solid])
