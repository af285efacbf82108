"""The known-answer cases of the reference's own test-suite, run against the CPU oracle.
Reference files: tests/kernels/{equilibrium,macroscopic,collision,stream}/test_*_jax.py,
tests/boundary_conditions/{bc_equilibrium,bc_fullway_bounce_back,mask}/test_*_jax.py."""

import numpy as np
import pytest

from oracle import lbm_numpy as O

LATTICES = [("D2Q9", (50, 50)), ("D3Q19", (30, 30, 30)), ("D3Q27", (30, 30, 30))]


def sphere_indices(shape):
    nr = shape[0]
    grids = np.meshgrid(*[np.arange(nr)] * len(shape))
    return np.array(np.where(sum((g - nr // 2) ** 2 for g in grids) < (nr // 4) ** 2))


@pytest.mark.parametrize("lattice,shape", LATTICES)
def test_equilibrium_of_rest_state_is_the_weights(lattice, shape):  # test_equilibrium_jax.py:30-48
    lat = O.Lattice(lattice)
    feq = O.equilibrium(np.ones((1,) + shape, np.float32), np.zeros((lat.d,) + shape, np.float32), lat)
    assert np.allclose(feq.sum(axis=0), 1.0)
    for l in range(lat.q):
        assert np.allclose(feq[l], lat.w[l])


@pytest.mark.parametrize("lattice,shape", LATTICES)
@pytest.mark.parametrize("rho,vel", [(1.0, 0.0), (1.1, 1.0), (1.1, 2.0)])
def test_macroscopic_inverts_equilibrium(lattice, shape, rho, vel):  # test_macroscopic_jax.py:23-50
    lat = O.Lattice(lattice)
    r = np.full((1,) + shape, rho, np.float32)
    u = np.full((lat.d,) + shape, vel, np.float32)
    r2, u2 = O.macroscopic(O.equilibrium(r, u, lat), lat)
    assert np.allclose(r2, rho) and np.allclose(u2, vel, atol=1e-6)


@pytest.mark.parametrize("lattice,shape", LATTICES)
@pytest.mark.parametrize("omega", [0.6, 1.0])
def test_bgk_identity(lattice, shape, omega):  # test_bgk_collision_jax.py:20-50
    lat = O.Lattice(lattice)
    feq = O.equilibrium(np.ones((1,) + shape, np.float32), np.zeros((lat.d,) + shape, np.float32), lat)
    f = np.zeros_like(feq)
    assert np.allclose(O.collide_bgk(f, feq, omega), f - omega * (f - feq))


@pytest.mark.parametrize("lattice,shape", LATTICES)
def test_stream_is_roll(lattice, shape):  # test_stream_jax.py:19-65: pins pull direction and periodic wrap
    lat = O.Lattice(lattice)
    f = np.zeros((lat.q,) + shape, np.float32)
    f[(slice(None),) + (slice(None),) * (lat.d - 1) + (shape[-1] // 2,)] = 1.0
    out = O.stream(f, lat)
    for l in range(lat.q):
        assert np.array_equal(out[l], np.roll(f[l], tuple(lat.c[:, l]), axis=tuple(range(lat.d))))


@pytest.mark.parametrize("lattice,shape", LATTICES)
@pytest.mark.parametrize("flavor", ["jax", "warp"])
def test_masker_ids(lattice, shape, flavor):  # test_bc_indices_masker_jax.py:31-79 (FullwayBB => no padding)
    lat = O.Lattice(lattice)
    idx = sphere_indices(shape)
    bc_mask, missing = O.build_masks([O.BC("fullway", 5, idx)], shape, lat, flavor=flavor)
    assert bc_mask.dtype == np.uint8 and missing.dtype == bool
    assert bc_mask.shape == (1,) + shape and missing.shape == (lat.q,) + shape
    assert np.all(bc_mask[(0,) + tuple(idx)] == 5)
    bc_mask[(0,) + tuple(idx)] = 0
    assert np.all(bc_mask == 0)


@pytest.mark.parametrize("lattice,shape", LATTICES)
def test_equilibrium_bc_on_sphere(lattice, shape):  # test_bc_equilibrium_jax.py:28-93
    lat = O.Lattice(lattice)
    idx = sphere_indices(shape)
    bc = O.BC("equilibrium", 1, idx, rho=1.0, u=(0.0,) * lat.d)
    bc_mask, missing = O.build_masks([bc], shape, lat, flavor="jax")
    f_pre = np.zeros((lat.q,) + shape, np.float32)
    f_post = np.full((lat.q,) + shape, 2.0, np.float32)
    f = O.apply_bc(bc, f_pre, f_post, bc_mask, missing, lat)
    outside = np.ones(shape, bool)
    outside[tuple(idx)] = False
    for l in range(lat.q):
        assert np.allclose(f[(l,) + tuple(idx)], lat.w[l])
        assert np.allclose(f[l][outside], 2.0)


@pytest.mark.parametrize("lattice,shape", LATTICES)
def test_fullway_bc_leaves_outside_untouched(lattice, shape):  # test_bc_fullway_bounce_back_jax.py:31-93
    lat = O.Lattice(lattice)
    idx = sphere_indices(shape)
    bc = O.BC("fullway", 1, idx)
    bc_mask, missing = O.build_masks([bc], shape, lat, flavor="jax")
    rng = np.random.default_rng(0)
    f_pre = rng.random((lat.q,) + shape, dtype=np.float32)
    f_post = np.full((lat.q,) + shape, 2.0, np.float32)
    f = O.apply_bc(bc, f_pre, f_post, bc_mask, missing, lat)
    outside = np.ones(shape, bool)
    outside[tuple(idx)] = False
    for l in range(lat.q):
        assert np.allclose(f[l][outside], 2.0)
        assert np.array_equal(f[(l,) + tuple(idx)], f_pre[(lat.opp[l],) + tuple(idx)])


def test_sharded_stream_equals_global_stream():  # distribute.py:23-44
    lat = O.Lattice("D3Q19")
    f = np.random.default_rng(3).random((19, 12, 5, 6)).astype(np.float32)
    for n in (1, 2, 3, 4):
        assert np.array_equal(O.stream_sharded(f, lat, n), O.stream(f, lat))


def test_mass_is_conserved_in_a_closed_periodic_box():
    lat = O.Lattice("D3Q19")
    rng = np.random.default_rng(0)
    shape = (10, 9, 8)
    f = O.initialize_eq(shape, lat, "FP64FP64", rho=1 + 1e-2 * rng.standard_normal((1,) + shape), u=1e-2 * rng.standard_normal((3,) + shape))
    m0 = f.sum()
    bm = np.zeros((1,) + shape, np.uint8)
    mm = np.zeros((19,) + shape, bool)
    f = O.run(f, bm, mm, [], 1.3, lat, 20, policy="FP64FP64")
    assert abs(f.sum() - m0) / m0 < 1e-13
