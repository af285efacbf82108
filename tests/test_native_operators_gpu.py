"""Stand-alone operators through the operator API on the GPU: (i) the reference test-suite's known answers, in the
reference's own style (WARP convention = buffers passed in, JAX convention = functional), (ii) seeded vectors produced
by the reference itself (tests/golden/operators.npz), (iii) both masker algorithms bit-exact against the oracle."""

import os

import numpy as np
import pytest
import torch

import xlb_b200 as xlb
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.grid import grid_factory
from xlb_b200.operator.boundary_masker import IndicesBoundaryMasker
from xlb_b200.operator.collision import BGK, KBC
from xlb_b200.operator.equilibrium import QuadraticEquilibrium
from xlb_b200.operator.macroscopic import Macroscopic, SecondMoment
from xlb_b200.operator.stream import Stream

from common import GOLDEN_DIR, rel_err

pytestmark = pytest.mark.gpu

CASES = [(2, "D2Q9", (50, 50)), (3, "D3Q19", (30, 30, 30)), (3, "D3Q27", (30, 30, 30))]


def init_xlb_env(lattice, backend=ComputeBackend.WARP, policy=xlb.PrecisionPolicy.FP32FP32):
    vel_set = getattr(xlb.velocity_set, lattice)(precision_policy=policy, compute_backend=backend)
    xlb.init(default_precision_policy=policy, default_backend=backend, velocity_set=vel_set)
    return vel_set


def sphere(grid_shape):
    nr = grid_shape[0]
    grids = np.meshgrid(*[np.arange(nr)] * len(grid_shape))
    idx = np.where(sum((g - nr // 2) ** 2 for g in grids) < (nr // 4) ** 2)
    return [tuple(int(v) for v in idx[i]) for i in range(len(grid_shape))]


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
def test_equilibrium_warp(dim, lattice, grid_shape):  # reference tests/kernels/equilibrium/test_equilibrium_warp.py
    vs = init_xlb_env(lattice)
    my_grid = grid_factory(grid_shape)
    rho = my_grid.create_field(cardinality=1, fill_value=1.0)
    u = my_grid.create_field(cardinality=dim, fill_value=0.0)
    f_eq = my_grid.create_field(cardinality=vs.q)
    f_eq = QuadraticEquilibrium()(rho, u, f_eq).numpy()
    assert np.allclose(f_eq.sum(axis=0), 1.0)
    for i, w in enumerate(vs.w):
        assert np.allclose(f_eq[i], w)


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
@pytest.mark.parametrize("rho,velocity", [(1.0, 0.0), (1.1, 1.0), (1.1, 2.0)])
@pytest.mark.parametrize("backend", [ComputeBackend.WARP, ComputeBackend.JAX])
def test_macroscopic(dim, lattice, grid_shape, rho, velocity, backend):  # tests/kernels/macroscopic/test_macroscopic_*.py
    vs = init_xlb_env(lattice, backend)
    my_grid = grid_factory(grid_shape)
    rho_field = my_grid.create_field(cardinality=1, fill_value=rho)
    velocity_field = my_grid.create_field(cardinality=dim, fill_value=velocity)
    if backend == ComputeBackend.WARP:
        f_eq = QuadraticEquilibrium()(rho_field, velocity_field, my_grid.create_field(cardinality=vs.q))
        rho_calc, u_calc = Macroscopic()(f_eq, my_grid.create_field(cardinality=1), my_grid.create_field(cardinality=dim))
    else:
        f_eq = QuadraticEquilibrium()(rho_field, velocity_field)
        rho_calc, u_calc = Macroscopic()(f_eq)
    assert np.allclose(rho_calc.numpy(), rho)
    assert np.allclose(u_calc.numpy(), velocity, atol=1e-06)


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
@pytest.mark.parametrize("omega", [0.6, 1.0])
def test_bgk_collision_warp(dim, lattice, grid_shape, omega):  # tests/kernels/collision/test_bgk_collision_warp.py
    vs = init_xlb_env(lattice)
    my_grid = grid_factory(grid_shape)
    rho = my_grid.create_field(cardinality=1, fill_value=1.0)
    u = my_grid.create_field(cardinality=dim, fill_value=0.0)
    f_eq = QuadraticEquilibrium()(rho, u, my_grid.create_field(cardinality=vs.q))
    f_orig = my_grid.create_field(cardinality=vs.q)
    f_out = my_grid.create_field(cardinality=vs.q)
    f_out = BGK()(f_orig, f_eq, f_out, rho, u, omega)
    f_eq, f_out, f_orig = f_eq.numpy(), f_out.numpy(), f_orig.numpy()
    assert np.allclose(f_out, f_orig - omega * (f_orig - f_eq), atol=1e-5)


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
@pytest.mark.parametrize("backend", [ComputeBackend.WARP, ComputeBackend.JAX])
def test_stream(dim, lattice, grid_shape, backend):  # tests/kernels/stream/test_stream_*.py: out[l] == roll(f[l], c_l)
    vs = init_xlb_env(lattice, backend)
    my_grid = grid_factory(grid_shape)
    f_initial = my_grid.create_field(cardinality=vs.q)
    f_np = np.zeros(tuple(f_initial.shape), dtype=np.float32)
    f_np[(slice(None),) + (slice(None),) * (dim - 1) + (grid_shape[-1] // 2,)] = 1.0
    f_np += np.random.default_rng(0).random(f_np.shape, dtype=np.float32)
    f_initial.copy_(torch.as_tensor(f_np))
    if backend == ComputeBackend.WARP:
        f_streamed = Stream()(f_initial, my_grid.create_field(cardinality=vs.q)).numpy()
    else:
        f_streamed = Stream()(f_initial).numpy()
    for i in range(vs.q):
        shift = tuple(int(s) for s in vs.c[:, i])
        assert np.array_equal(f_streamed[i], np.roll(f_np[i], shift, axis=tuple(range(dim))))


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
@pytest.mark.parametrize("backend", [ComputeBackend.WARP, ComputeBackend.JAX])
def test_indices_masker(dim, lattice, grid_shape, backend):  # tests/boundary_conditions/mask/test_bc_indices_masker_*.py
    vs = init_xlb_env(lattice, backend)
    my_grid = grid_factory(grid_shape)
    missing_mask = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    bc_mask = my_grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    indices = sphere(grid_shape)
    test_bc = xlb.operator.boundary_condition.FullwayBounceBackBC(indices=indices)
    test_bc.id = 5
    bc_mask, missing_mask = IndicesBoundaryMasker()([test_bc], bc_mask, missing_mask)
    assert missing_mask.dtype == xlb.Precision.BOOL.wp_dtype
    assert bc_mask.dtype == xlb.Precision.UINT8.wp_dtype
    bc_mask = bc_mask.numpy()
    if dim == 2 and backend == ComputeBackend.WARP:
        assert bc_mask.shape == (1,) + grid_shape + (1,)
        bc_mask = bc_mask[..., 0]
    else:
        assert bc_mask.shape == (1,) + grid_shape
    assert np.all(bc_mask[(0,) + tuple(indices)] == test_bc.id)
    bc_mask[(0,) + tuple(indices)] = 0
    assert np.all(bc_mask == 0)


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
def test_bc_equilibrium_warp(dim, lattice, grid_shape):  # tests/boundary_conditions/bc_equilibrium/test_bc_equilibrium_warp.py
    vs = init_xlb_env(lattice)
    my_grid = grid_factory(grid_shape)
    missing_mask = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    bc_mask = my_grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    indices = sphere(grid_shape)
    equilibrium_bc = xlb.operator.boundary_condition.EquilibriumBC(
        rho=1.0, u=(0.0, 0.0, 0.0) if dim == 3 else (0.0, 0.0), equilibrium_operator=QuadraticEquilibrium(), indices=indices
    )
    bc_mask, missing_mask = IndicesBoundaryMasker()([equilibrium_bc], bc_mask, missing_mask, start_index=None)
    f_pre = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.FP32)
    f_post = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.FP32, fill_value=2.0)
    f = equilibrium_bc(f_pre, f_post, bc_mask, missing_mask).numpy()
    if dim == 2:
        assert f.shape == (vs.q,) + grid_shape + (1,)
        f = f[..., 0]
    outside = np.ones(grid_shape, dtype=bool)
    outside[tuple(indices)] = False
    for i, w in enumerate(vs.w):
        assert np.allclose(f[(i,) + tuple(indices)], w)
        assert np.allclose(f[i][outside], 2.0)


@pytest.mark.parametrize("dim,lattice,grid_shape", CASES)
@pytest.mark.parametrize("backend", [ComputeBackend.WARP, ComputeBackend.JAX])
def test_bc_fullway_bounce_back(dim, lattice, grid_shape, backend):  # tests/boundary_conditions/bc_fullway_bounce_back/*
    vs = init_xlb_env(lattice, backend)
    my_grid = grid_factory(grid_shape)
    missing_mask = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    bc_mask = my_grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    indices = sphere(grid_shape)
    bc = xlb.operator.boundary_condition.FullwayBounceBackBC(indices=indices)
    bc_mask, missing_mask = IndicesBoundaryMasker()([bc], bc_mask, missing_mask, start_index=None)
    f_pre = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.FP32)
    f_pre.copy_(torch.rand(tuple(f_pre.shape)))
    f_post = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.FP32, fill_value=2.0)
    f = bc(f_pre, f_post, bc_mask, missing_mask).numpy().reshape((vs.q,) + grid_shape)
    pre = f_pre.numpy().reshape((vs.q,) + grid_shape)
    outside = np.ones(grid_shape, dtype=bool)
    outside[tuple(indices)] = False
    for i in range(vs.q):
        assert np.allclose(f[i][outside], 2.0)
        assert np.array_equal(f[(i,) + tuple(indices)], pre[(int(vs.opp_indices[i]),) + tuple(indices)])


@pytest.mark.parametrize("lattice", ["D2Q9", "D3Q19", "D3Q27"])
def test_operators_match_reference_vectors(lattice):
    """Seeded single-operator vectors computed by the reference's own JAX implementations."""
    z = np.load(os.path.join(GOLDEN_DIR, "operators.npz"))
    vs = init_xlb_env(lattice, ComputeBackend.JAX)
    f, rho, u = z[f"{lattice}_f"], z[f"{lattice}_rho"], z[f"{lattice}_u"]
    assert np.array_equal(Stream()(f).numpy(), z[f"{lattice}_stream"])
    assert rel_err(QuadraticEquilibrium()(rho, u).numpy(), z[f"{lattice}_feq"]) <= 1e-6
    r2, u2 = Macroscopic()(f)
    assert rel_err(r2.numpy(), z[f"{lattice}_rho2"]) <= 1e-6 and np.abs(u2.numpy() - z[f"{lattice}_u2"]).max() <= 1e-6
    assert np.abs(SecondMoment()(f).numpy() - z[f"{lattice}_pi"]).max() <= 1e-5
    feq2 = QuadraticEquilibrium()(r2, u2)
    assert rel_err(BGK()(f, feq2, r2, u2, 1.3).numpy(), z[f"{lattice}_bgk"]) <= 1e-6
    if lattice != "D3Q19":
        assert rel_err(KBC()(f, feq2, r2, u2, 1.7).numpy(), z[f"{lattice}_kbc"]) <= 1e-5
    else:
        with pytest.raises(NotImplementedError):
            KBC()


@pytest.mark.parametrize("flavor,backend", [("warp", ComputeBackend.WARP), ("jax", ComputeBackend.JAX)])
@pytest.mark.parametrize("lattice,shape", [("D2Q9", (40, 36)), ("D3Q19", (24, 20, 16)), ("D3Q27", (24, 20, 16))])
def test_both_masker_algorithms_bit_exact(flavor, backend, lattice, shape):
    """Interior HalfwayBB sphere (needs_padding) + face BCs + a BC-free periodic pair of faces: the case where the
    reference's two algorithms DIFFER; each must match its oracle restatement bit for bit."""
    from oracle import lbm_numpy as O
    from xlb_b200.operator.boundary_condition import FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC

    vs = init_xlb_env(lattice, backend)
    lat = O.Lattice(lattice)
    d = lat.d
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    sph = np.array(np.where(sum((g - s // 3) ** 2 for g, s in zip(grids, shape)) < (min(shape) // 5) ** 2))
    box = O.bounding_box_indices(shape, remove_edges=True)
    walls = np.concatenate([box["bottom"], box["top"]], axis=1)
    inlet = box["left"]
    obcs = [O.BC("fullway", 1, walls), O.BC("regularized", 2, inlet), O.BC("halfway", 3, sph)]
    bm, mm = O.build_masks(obcs, shape, lat, flavor=flavor)

    my_grid = grid_factory(shape)
    bcs = [
        FullwayBounceBackBC(indices=walls.tolist()),
        RegularizedBC("velocity", prescribed_value=(0.01,) + (0.0,) * (d - 1), indices=inlet.tolist()),
        HalfwayBounceBackBC(indices=sph.tolist()),
    ]
    missing_mask = my_grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    bc_mask = my_grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    bc_mask, missing_mask = IndicesBoundaryMasker()(bcs, bc_mask, missing_mask)
    assert np.array_equal(bc_mask.numpy().reshape(bm.shape), bm)
    assert np.array_equal(missing_mask.numpy().reshape(mm.shape), mm)


@pytest.mark.parametrize("backend", [ComputeBackend.WARP, ComputeBackend.JAX])
@pytest.mark.parametrize("name", ["sphere_d3q27_kbc_fp32", "sphere_d3q19_bgk_fp32", "sphere_d3q27_bgk_regpressure_fp64"])
def test_momentum_transfer_matches_reference_vectors(name, backend):
    """Drag / lift of the sphere by momentum exchange, against the value the reference's MomentumTransfer computed."""
    from common import load_golden, native_case
    from xlb_b200.operator.force import MomentumTransfer

    g = load_golden(name)
    stepper, f_0, f_1, bc_mask, missing_mask = native_case(g, backend=backend.name)
    f_0.copy_(torch.as_tensor(g["f_final"]).reshape(f_0.shape))
    bc = stepper.boundary_conditions[int(g["force_bc"])]
    force = MomentumTransfer(bc)(f_0, f_1, bc_mask, missing_mask)
    force = force.numpy() if hasattr(force, "numpy") and not isinstance(force, np.ndarray) else np.asarray(force)
    assert force.shape == (3,)
    assert np.allclose(force, g["force"], rtol=2e-5, atol=1e-7), (force, g["force"])


@pytest.mark.parametrize("lattice", ["D3Q19", "D3Q27", "D2Q9"])
def test_extended_collision_operators_standalone(lattice):
    """SmagorinskyLESBGK / ForcedCollision / ExactDifference through the operator classes (xlbn_collide_ext, xlbn_exact_difference;
    reference: smagorinsky_les_bgk.py:92-138, forced_collision.py:46-50, exact_difference_force.py:79-84) vs the numpy oracle on a
    seeded random state.  Two gates per operator: the full populations at 5e-7 relative (a few fp32 ulps of f), and the
    operator's INCREMENT (out - f) at the same absolute bound — the ExactDifference increment itself is only ~3e-5 |f|, so a
    bound relative to the increment would sit below one ulp of the populations it is added to (round 1's mistake)."""
    from oracle import lbm_numpy as O
    from xlb_b200.operator.collision import ForcedCollision, SmagorinskyLESBGK
    from xlb_b200.operator.force import ExactDifference

    vs = init_xlb_env(lattice)
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
    scale = np.abs(f).max()

    def check(name, got, want):
        got = back(got)
        assert rel_err(got, want) <= 5e-7, f"{lattice} {name}: full rel err {rel_err(got, want):.2e}"
        assert np.abs((got - f) - (want - f)).max() <= 5e-7 * scale, f"{lattice} {name}: increment"
        assert np.abs(want - f).max() > 0, name  # the operator did something

    check("ExactDifference", ExactDifference(force)(F, FEQ, torch.empty_like(F), RHO, U), O.exact_difference_force(f.copy(), feq, rho, u, force, lat))
    cases = [("BGK", BGK)] + ([("KBC", KBC)] if lattice != "D3Q19" else []) + ([("SmagorinskyLESBGK", SmagorinskyLESBGK)] if lat.d == 3 else [])
    for cname, cls in cases:
        if cname == "BGK":
            base = O.collide_bgk(f, feq, 1.7)
        elif cname == "KBC":
            base = O.collide_kbc(f, feq, rho, lat, 1.7)
        else:
            base = O.collide_smagorinsky(f, feq, lat, 1.7)
            check(cname, cls()(F, FEQ, RHO, U, torch.empty_like(F), 1.7), base)
        want = O.exact_difference_force(base, feq, rho, u, force, lat)
        check("Forced" + cname, ForcedCollision(cls(), force_vector=force)(F, FEQ, torch.empty_like(F), RHO, U, 1.7), want)
