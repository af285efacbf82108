"""Parity of the fused CUDA step (through the operator API -> C ABI) against (a) vectors produced by the reference's own
Python and (b) the CPU oracle.  Masks bit-exact; populations within the north-star tolerances
(1e-5 relative fp32, 1e-3 fp16 storage)."""

import numpy as np
import pytest

from common import RTOL, STEP_CASES, load_golden, native_run, oracle_run, rel_err, unpack_bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["WARP", "JAX"])
@pytest.mark.parametrize("name", STEP_CASES)
def test_step_matches_reference_vectors(name, backend):
    g = load_golden(name)
    q = g["f_final"].shape[0]
    f, bc_mask, missing = native_run(g, backend=backend)
    if g["n_bc"]:
        if backend == "JAX" and len(g["shape"]) == 2:
            pass
        assert np.array_equal(bc_mask.reshape(g["bc_mask"].shape), g["bc_mask"]), "bc_mask must be bit-exact"
        assert np.array_equal(missing.reshape((q,) + g["shape"]), unpack_bits(g["missing_bits"], q)), "missing_mask must be bit-exact"
    assert f.dtype == g["f_final"].dtype
    assert np.isfinite(f.astype(np.float64)).all()
    assert rel_err(f, g["f_final"]) <= RTOL[g["policy"]], f"{name}: rel err {rel_err(f, g['f_final']):.3e}"


@pytest.mark.parametrize("v", [1, 2, 4, 8, 102, 104, 202])  # 10x = packed fp32x2 pair path, 202 = half2-state path (FP32FP16 BGK)
@pytest.mark.parametrize("name", ["cavity_d3q19_bgk_fp32", "cavity_d3q19_bgk_fp32fp16", "sphere_d3q27_kbc_fp32", "sphere_d3q27_bgk_regpressure_fp64", "cavity_d2q9_kbc_fp32"])
def test_every_cells_per_thread_variant(name, v):
    """All vector widths of the kernel compute the same step (the library falls back when v does not divide nz)."""
    g = load_golden(name)
    if v == 202 and not (g["policy"] == "FP32FP16" and g["collision"] == "BGK"):
        with pytest.raises(Exception, match="FP32FP16 BGK only"):
            native_run(g, cells_per_thread=v)
        return
    if v >= 100 and g["policy"].startswith("FP64"):
        with pytest.raises(Exception, match="packed pair path"):
            native_run(g, cells_per_thread=v)
        return
    f, _, _ = native_run(g, cells_per_thread=v)
    assert rel_err(f, g["f_final"]) <= RTOL[g["policy"]]


def test_cavity_1000_steps_against_oracle():
    """C1-style run (examples/performance/mlups_3d.py set-up) for 1000 steps at a size the numpy oracle finishes in
    seconds; tolerance = north-star 1e-5 relative on populations, rho and u."""
    from oracle import lbm_numpy as O

    g = load_golden("cavity_d3q19_bgk_fp32")
    n = 32
    lat = O.Lattice("D3Q19")
    box, box_ne = O.bounding_box_indices((n,) * 3), O.bounding_box_indices((n,) * 3, remove_edges=True)
    walls = np.unique(np.concatenate([box[k] for k in ("bottom", "left", "right", "front", "back")], axis=1), axis=-1)
    g.update(shape=(n, n, n), steps=1000, omega=1.0, f_init=O.initialize_eq((n,) * 3, lat))
    g["bcs"] = [dict(kind="equilibrium", id=1, indices=box_ne["top"], rho=1.0, u=np.array([0.02, 0, 0])), dict(kind="fullway", id=2, indices=walls)]
    ref, bm, mm = oracle_run(g)
    f, bc_mask, missing = native_run(g)
    assert np.array_equal(bc_mask, bm) and np.array_equal(missing, mm)
    assert rel_err(f, ref) <= 1e-5
    rho_r, u_r = O.macroscopic(ref, lat)
    rho_n, u_n = O.macroscopic(f, lat)
    assert rel_err(rho_n, rho_r) <= 1e-5
    assert np.abs(u_n - u_r).max() <= 1e-5 * max(np.abs(u_r).max(), 0.02)


def test_solid_cells_255_are_skipped():
    """bc_mask == 255 => the cell is neither read nor written (reference: nse_stepper.py:356-358)."""
    import torch

    g = load_golden("cavity_d3q19_bgk_fp32")
    from common import native_case

    stepper, f_0, f_1, bc_mask, missing_mask = native_case(g)
    bc_mask[0, 5:9, 5:9, 4:12] = 255
    f_1.fill_(7.0)
    before = f_1.clone()
    stepper(f_0, f_1, bc_mask, missing_mask, 1.0, 0)
    solid = (bc_mask[0] == 255).unsqueeze(0).expand_as(f_1)
    assert torch.equal(f_1[solid], before[solid])
    assert not torch.equal(f_1[~solid], before[~solid])


def test_stepper_errors_are_loud():
    import torch

    from common import native_case

    g = load_golden("cavity_d3q19_bgk_fp32")
    stepper, f_0, f_1, bc_mask, missing_mask = native_case(g)
    with pytest.raises(Exception, match="no CPU fallback"):
        stepper(f_0.cpu(), f_1, bc_mask, missing_mask, 1.0, 0)
    with pytest.raises(Exception, match="different buffers"):
        stepper(f_0, f_0, bc_mask, missing_mask, 1.0, 0)
    with pytest.raises(Exception):
        stepper(f_0.to(torch.float16), f_1, bc_mask, missing_mask, 1.0, 0)


@pytest.mark.parametrize("v", [1, 102])
def test_c3_sphere_d3q27_kbc_256x64x64_against_c_oracle(v):
    """BASELINE config C3 at the reference example's own size (examples/cfd/flow_past_sphere_3d.py:22: 256x64x64):
    D3Q27 KBC omega 1.6, Fullway walls, Regularized Poiseuille inlet, ExtrapolationOutflow outlet, Halfway sphere.
    300 steps against the C oracle; masks bit-exact, populations within 1e-5 relative."""
    from oracle import lbm_c
    from oracle import lbm_numpy as O

    if not lbm_c.available():
        pytest.skip("oracle/liblbm_ref.so not built")
    shape, lat = (256, 64, 64), O.Lattice("D3Q27")
    box, bne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
    walls = np.unique(np.concatenate([box[k] for k in ("bottom", "top", "front", "back")], axis=1), axis=-1)
    X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    sph = np.array(np.where((X - shape[0] // 6) ** 2 + (Y - shape[1] // 2) ** 2 + (Z - shape[2] // 2) ** 2 < (shape[1] // 12) ** 2))
    Hy, Hz = float(shape[1] - 1), float(shape[2] - 1)
    yy, zz = np.meshgrid(np.arange(shape[1]), np.arange(shape[2]), indexing="ij")
    ux = (0.04 * np.maximum(0.0, 1.0 - ((2.0 * (yy - Hy / 2.0) / Hy) ** 2.0 + (2.0 * (zz - Hz / 2.0) / Hz) ** 2.0))).astype(np.float32)
    pv = np.stack([ux, np.zeros_like(ux), np.zeros_like(ux)])
    g = load_golden("sphere_d3q27_kbc_fp32")
    g.update(shape=shape, steps=300, omega=1.6, f_init=O.initialize_eq(shape, lat))
    g["bcs"] = [
        dict(kind="fullway", id=1, indices=walls),
        dict(kind="regularized", id=2, indices=bne["left"], bc_type="velocity", prescribed=pv),
        dict(kind="outflow", id=3, indices=bne["right"]),
        dict(kind="halfway", id=4, indices=sph),
    ]
    from common import oracle_bcs

    bcs = oracle_bcs(g)
    bc_mask, missing = O.build_masks(bcs, shape, lat, flavor="warp")
    ref = lbm_c.run(g["f_init"], bc_mask, missing, bcs, 1.6, lat, 300, "FP32FP32", "KBC")
    f, bm, mm = native_run(g, cells_per_thread=v)
    assert np.array_equal(bm, bc_mask) and np.array_equal(mm, missing)
    assert np.isfinite(f).all()
    assert rel_err(f, ref) <= 1e-5, rel_err(f, ref)


@pytest.mark.parametrize("policy", ["FP32FP32", "FP32FP16"])
@pytest.mark.parametrize("shape", [(10, 9, 7), (6, 5, 9)])
def test_odd_extents_fall_back_to_one_cell_per_thread(policy, shape):
    """Odd nz cannot use the vector / pair variants: the library falls back to the one-cell path and stays correct."""
    from oracle import lbm_numpy as O

    lat = O.Lattice("D3Q19")
    rng = np.random.default_rng(5)
    cdt, sdt = O.policy_dtypes(policy)
    f_init = O.initialize_eq(shape, lat, policy, rho=1 + 1e-2 * rng.standard_normal((1,) + shape), u=1e-2 * rng.standard_normal((3,) + shape))
    g = load_golden("periodic_d3q19_bgk_fp32")
    g.update(shape=shape, steps=10, omega=1.4, policy=policy, f_init=f_init, bcs=[], n_bc=0)
    ref, _, _ = oracle_run(g)
    for v in (0, 2, 4, 202 if policy == "FP32FP16" else 102):
        f, _, _ = native_run(g, cells_per_thread=v)
        assert rel_err(f, ref) <= RTOL[policy]
