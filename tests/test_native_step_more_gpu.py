"""GPU parity of the fused step, second file: cases that first ran on a B200 in round 2 (they were `xfail`-wrapped first-run
cases at the end of round 1; all are ordinary hard tests now).

  * 2-D channels and the non-trivial BCs under the other precision policies (reference JAX-backend vectors),
  * every vector produced by the reference's WARP backend (tests/golden/make_golden_warp.py), incl. FP32FP16 ones whose
    prescribed inlet value is rounded to the store dtype exactly as the Warp path does (boundary_condition.py:151),
  * the extended collision operators in the fused step (SmagorinskyLESBGK, ForcedCollision: SURVEY §8f N4),
  * the KBC formulations (register-lean = default, literal = cells_per_thread 300) and the half2-state variants,
  * stepper.run (CUDA graph replay) == n individual calls, bit for bit.
Tolerances: tests/common.py RTOL (north-star: 1e-5 relative fp32, 1e-3 fp16 storage); masks bit-exact."""

import numpy as np
import pytest
import torch

from common import LATE_CASES, RTOL, STEP_CASES, WARP_CASES, WARP_CASES_FP16, WARP_CASES_N4, load_golden, native_case, native_run, rel_err, rel_err_elem, unpack_bits

pytestmark = pytest.mark.gpu
KBC_CASES = [n for n in STEP_CASES + WARP_CASES + WARP_CASES_FP16 if "kbc" in n]


def check_step_case(name, backend, v=0):
    g = load_golden(name)
    q = g["f_final"].shape[0]
    f, bc_mask, missing = native_run(g, backend=backend, cells_per_thread=v)
    assert np.array_equal(bc_mask.reshape(g["bc_mask"].shape), g["bc_mask"]), "bc_mask must be bit-exact"
    assert np.array_equal(missing.reshape((q,) + g["shape"]), unpack_bits(g["missing_bits"], q)), "missing_mask must be bit-exact"
    err = rel_err(f, g["f_final"])
    assert err <= RTOL[g["policy"]], f"{name} {backend} v={v}: rel err {err:.3e} (element-relative {rel_err_elem(f, g['f_final']):.3e})"
    if "force" in g and len(g["shape"]) == 3:
        # MomentumTransfer on the reference's final populations: same input, so only the summation arithmetic differs; a vector
        # stored in fp16 (FP32FP16 policy) carries the reference's own fp16 rounding of the three components
        from xlb_b200.operator.force import MomentumTransfer

        stepper, f_0, f_1, bm, mm = native_case(g, backend=backend)
        f_0.copy_(torch.as_tensor(g["f_final"]).reshape(f_0.shape))
        force = MomentumTransfer(stepper.boundary_conditions[int(g["force_bc"])])(f_0, f_1, bm, mm)
        force = np.asarray(force.numpy() if hasattr(force, "numpy") and not isinstance(force, np.ndarray) else force, dtype=np.float64)
        want = np.asarray(g["force"], dtype=np.float64)
        rtol = 2e-3 if g["force"].dtype == np.float16 else 2e-5
        assert np.allclose(force, want, rtol=rtol, atol=rtol * 0.1 * np.abs(want).max()), f"{name}: force {force} vs {want}"


@pytest.mark.parametrize("backend", ["WARP", "JAX"])
@pytest.mark.parametrize("name", LATE_CASES)
def test_late_reference_vectors(name, backend):
    check_step_case(name, backend)


@pytest.mark.parametrize("name", WARP_CASES + WARP_CASES_FP16)
def test_reference_warp_backend_vectors(name):
    """Vectors from the reference's own WARP backend: 255 skip, scalar prescribed value in f_1[0] (store dtype), per-index interior
    flag of the Warp masker, Warp outflow neighbour read."""
    check_step_case(name, "WARP")


@pytest.mark.parametrize("v", [1, 2, 202, 203])
def test_fp16_sphere_every_path_against_the_warp_convention_vector(v):
    """FP32FP16 with Regularized inlet + ExtrapolationOutflow + Halfway body (the configuration that was red in round 1 — because
    of the test's force tolerance, not the kernel): every code path vs the WARP-backend vector of the same case."""
    check_step_case("warp_sphere_d3q19_bgk_fp32fp16", "WARP", v)
    check_step_case("sphere_d3q19_bgk_fp32fp16", "WARP", v)  # JAX-convention vector: differs by the fp16 rounding of the inlet value (7e-4)


@pytest.mark.parametrize("name", WARP_CASES_N4)
def test_extended_collision_kernels(name):
    check_step_case(name, "WARP")


@pytest.mark.parametrize("v", [300, 301, 1])
@pytest.mark.parametrize("name", KBC_CASES)
def test_kbc_formulations(name, v):
    """301 = register-lean KBC (the default since round 2), 300 = the literal three-array formulation, 1 = literal, scalar path."""
    check_step_case(name, "WARP", v)


@pytest.mark.parametrize("name,n", [("cavity_d3q19_bgk_fp32", 10), ("cavity_d3q19_bgk_fp32fp16", 10), ("sphere_d3q27_kbc_fp32", 12), ("cavity_d2q9_kbc_fp32", 7)])
def test_cuda_graph_loop_equals_individual_calls(name, n):
    """stepper.run(n): a captured pair of steps replayed n/2 times must give the bits of n individual calls."""
    g = load_golden(name)
    stepper, f_0, f_1, bm, mm = native_case(g)
    a, b = stepper.run(f_0, f_1, bm, mm, g["omega"], n)
    stepper2, g_0, g_1, bm2, mm2 = native_case(g)
    for rounds in range(2):  # second call: same buffers (n even) -> replay only
        if rounds:
            a, b = stepper.run(a, b, bm, mm, g["omega"], n)
        for i in range(n):
            g_0, g_1 = stepper2(g_0, g_1, bm2, mm2, g["omega"], i)
            g_0, g_1 = g_1, g_0
        assert torch.equal(a, g_0)


def test_omega_may_change_every_step_without_a_host_sync():
    """omega is a per-call argument of the reference kernel (nse_stepper.py:351).  Under FP32FP16 the EquilibriumBC constants follow it
    stream-ordered: a ramp gives the same bits as fresh steppers created per omega, on the pair path and on the scalar path."""
    g = load_golden("cavity_d3q19_bgk_fp32fp16")
    omegas = [1.0 + 0.05 * i for i in range(8)]
    out = {}
    for v in (202, 1):
        stepper, f_0, f_1, bm, mm = native_case(g, cells_per_thread=v)
        for i, om in enumerate(omegas):
            f_0, f_1 = stepper(f_0, f_1, bm, mm, om, i)
            f_0, f_1 = f_1, f_0
        out[v] = f_0.clone()
    assert torch.equal(out[202], out[1])
