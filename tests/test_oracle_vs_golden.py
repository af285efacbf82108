"""Pins the CPU oracle (oracle/lbm_numpy.py) against vectors produced by the reference's own Python
(tests/golden/make_golden.py).  CPU only."""

import numpy as np
import pytest

from common import LATE_CASES, STEP_CASES, load_golden, oracle_run, rel_err, unpack_bits
from oracle import lbm_numpy as O


@pytest.mark.parametrize("name", STEP_CASES + LATE_CASES)
def test_step_cases(name):
    g = load_golden(name)
    lat = O.Lattice(g["lattice"])
    f, bc_mask, missing = oracle_run(g, flavor="jax")
    # masks: bit-exact (integer / bool work)
    assert np.array_equal(bc_mask, g["bc_mask"])
    assert np.array_equal(missing, unpack_bits(g["missing_bits"], lat.q))
    # populations: same algorithm in the same dtype on the same numpy -> identical to the last bit
    assert f.dtype == g["f_final"].dtype
    assert rel_err(f, g["f_final"]) <= 1e-7
    rho, u = O.macroscopic(f.astype(g["rho"].dtype), lat)
    assert rel_err(rho, g["rho"]) <= 1e-6 and np.abs(u - g["u"]).max() <= 1e-6


@pytest.mark.parametrize("name", [c for c in STEP_CASES if not c.startswith("periodic")])
def test_warp_masker_equals_jax_masker_on_bounded_domains(name):
    """The reference's two masker algorithms agree on all configs whose domain faces carry BCs."""
    g = load_golden(name)
    lat = O.Lattice(g["lattice"])
    from common import oracle_bcs

    bm_w, mm_w = O.build_masks(oracle_bcs(g), g["shape"], lat, flavor="warp")
    assert np.array_equal(bm_w, g["bc_mask"])
    assert np.array_equal(mm_w, unpack_bits(g["missing_bits"], lat.q))


@pytest.mark.parametrize("lattice", ["D2Q9", "D3Q19", "D3Q27"])
def test_operator_vectors(lattice):
    z = np.load(__import__("os").path.join(__import__("common").GOLDEN_DIR, "operators.npz"))
    lat = O.Lattice(lattice)
    f, rho, u = z[f"{lattice}_f"], z[f"{lattice}_rho"], z[f"{lattice}_u"]
    assert np.array_equal(O.stream(f, lat), z[f"{lattice}_stream"])
    assert rel_err(O.equilibrium(rho, u, lat), z[f"{lattice}_feq"]) <= 1e-7
    r2, u2 = O.macroscopic(f, lat)
    assert rel_err(r2, z[f"{lattice}_rho2"]) <= 1e-7 and np.abs(u2 - z[f"{lattice}_u2"]).max() <= 1e-7
    assert np.abs(O.second_moment(f, lat) - z[f"{lattice}_pi"]).max() <= 1e-6
    feq2 = O.equilibrium(r2, u2, lat)
    assert rel_err(O.collide_bgk(f, feq2, 1.3), z[f"{lattice}_bgk"]) <= 1e-7
    if lattice != "D3Q19":
        assert rel_err(O.collide_kbc(f, feq2, r2, lat, 1.7), z[f"{lattice}_kbc"]) <= 1e-6
    else:
        with pytest.raises(NotImplementedError):
            O.collide_kbc(f, feq2, r2, lat, 1.7)


@pytest.mark.parametrize("name", [c for c in STEP_CASES if c.startswith("sphere")])
def test_momentum_transfer_vectors(name):
    """MomentumTransfer (operator/force/momentum_transfer.py:51-90) on the final state of the sphere cases."""
    from common import oracle_bcs

    g = load_golden(name)
    lat = O.Lattice(g["lattice"])
    cdt, _ = O.policy_dtypes(g["policy"])
    bc = oracle_bcs(g)[int(g["force_bc"])]
    force = O.momentum_transfer(bc, g["f_final"].astype(cdt), g["bc_mask"], unpack_bits(g["missing_bits"], lat.q), lat)
    assert np.allclose(force, g["force"], rtol=1e-6, atol=1e-9)


def test_forced_collision_vector():
    """ForcedCollision + ExactDifference (collision/forced_collision.py:34-39, force/exact_difference_force.py:45-70):
    oracle vs the reference's own run.  The CUDA path does not implement body forces yet."""
    g = load_golden("periodic_d3q19_bgk_forced_fp32")
    lat = O.Lattice(g["lattice"])
    bm = np.zeros((1,) + g["shape"], np.uint8)
    mm = np.zeros((lat.q,) + g["shape"], bool)
    f = O.run(g["f_init"].copy(), bm, mm, [], g["omega"], lat, g["steps"], policy=g["policy"], collision=g["collision"], force=g["force_vector"])
    assert rel_err(f, g["f_final"]) <= 1e-7
    unforced = O.run(g["f_init"].copy(), bm, mm, [], g["omega"], lat, g["steps"], policy=g["policy"], collision=g["collision"])
    assert rel_err(unforced, g["f_final"]) > 1e-6  # the force does something
