"""x-slab halo on ONE device: N slabs of a domain live on the same GPU, connected through the same ghost-plane /
counter protocol the multi-GPU path uses (xlbn_halo_*, peer pointers = plain device pointers).  The decomposed run must
be BIT-IDENTICAL to the undecomposed one (same arithmetic per cell)."""

import ctypes as C

import numpy as np
import pytest
import torch

from common import load_golden, native_case
from xlb_b200 import native

pytestmark = pytest.mark.gpu


def run_slabs(g, n_slabs, steps, split_faces):
    stepper, f_0, f_1, bc_mask, missing_mask = native_case(g)
    handle = stepper._native_handle()
    bits_all = stepper._missing_bits(missing_mask).reshape(g["shape"]) if stepper._needs_missing else None
    L = native.lib()
    nx, ny, nz = tuple(g["shape"]) + (1,) * (3 - len(g["shape"]))  # 2-D fields: [q][nx][ny][1]
    assert nx % n_slabs == 0
    nxl = nx // n_slabs
    st = native.stream_of(f_0)

    def slab(t, i):
        return t[:, i * nxl : (i + 1) * nxl].contiguous()

    F0 = [slab(f_0, i) for i in range(n_slabs)]
    F1 = [slab(f_1, i) for i in range(n_slabs)]
    BC = [slab(bc_mask, i) for i in range(n_slabs)]
    BITS = [bits_all[i * nxl : (i + 1) * nxl].contiguous() if bits_all is not None else None for i in range(n_slabs)]
    halos = []
    for i in range(n_slabs):
        h = C.c_void_p()
        native.check(L.xlbn_halo_create(stepper._lattice, stepper.precision_policy.store_precision.code, ny, nz, C.byref(h)))
        halos.append(h)
    bases = []
    for h in halos:
        p, nbytes = C.c_void_p(), C.c_longlong()
        native.check(L.xlbn_halo_ghost_ptr(h, C.byref(p), C.byref(nbytes)))
        bases.append(p.value)
    for i, h in enumerate(halos):
        lo = C.create_string_buffer(np.uint64(bases[(i - 1) % n_slabs]).tobytes(), 64)
        hi = C.create_string_buffer(np.uint64(bases[(i + 1) % n_slabs]).tobytes(), 64)
        native.check(L.xlbn_halo_connect(h, lo, hi, 1))
    full = native.Domain(nxl, ny, nz, 0, nxl)
    for i, h in enumerate(halos):
        native.check(L.xlbn_halo_push(h, native.ptr(F0[i]), C.byref(full), 1, st))
        native.check(L.xlbn_halo_push(h, native.ptr(F0[i]), C.byref(full), 0, st))
        native.check(L.xlbn_halo_signal(h, 0, st))
    for t in range(steps):
        for i, h in enumerate(halos):
            args = (handle, native.ptr(F0[i]), native.ptr(F1[i]), native.ptr(BC[i]), native.ptr(BITS[i]))
            native.check(L.xlbn_halo_wait(h, t, st))
            if split_faces and nxl >= 3:
                for x0, cnt, hh in ((0, 1, h), (nxl - 1, 1, h), (1, nxl - 2, None)):
                    dom = native.Domain(nxl, ny, nz, x0, cnt)
                    native.check(L.xlbn_step(*args, C.byref(dom), g["omega"], t, hh, st))
            else:
                native.check(L.xlbn_step(*args, C.byref(full), g["omega"], t, h, st))
            native.check(L.xlbn_halo_signal(h, t + 1, st))
        F0, F1 = F1, F0
    torch.cuda.synchronize()
    out = torch.cat(F0, dim=1).cpu().numpy()
    if len(g["shape"]) == 2:
        out = out[..., 0]
    for h in halos:
        L.xlbn_halo_destroy(h)
    return out


@pytest.mark.parametrize("n_slabs", [2, 4])
@pytest.mark.parametrize("split_faces", [False, True])
@pytest.mark.parametrize("name", ["cavity_d3q19_bgk_fp32", "sphere_d3q27_kbc_fp32", "sphere_d3q19_bgk_zouhe_pressure_fp32", "periodic_d3q19_bgk_fp32", "cavity_d3q19_bgk_fp32fp16"])
def test_slab_run_is_bit_identical(name, n_slabs, split_faces):
    from common import native_run

    g = load_golden(name)
    steps = 12
    whole, _, _ = native_run(g, steps=steps)
    parts = run_slabs(g, n_slabs, steps, split_faces)
    assert np.array_equal(parts, whole)


@pytest.mark.parametrize("n_slabs", [2, 4])
@pytest.mark.parametrize("policy", ["FP32FP32", "FP64FP32", "FP32FP16"])
@pytest.mark.parametrize("walls", [True, False])
def test_slab_run_with_tile_kernel_interiors_is_bit_identical(policy, n_slabs, walls):
    """Shapes the tile kernels accept (csrc/step_tile.cuh): the interior launch of every slab (no halo handle on that call) takes the TMA-fed
    tile kernel on the partial x range [1, nxl - 1), the two face planes the direct kernel with peer stores; together they must give the
    bits of the undecomposed run (which tiles the whole box)."""
    from common import native_run, tile_case

    g = tile_case("D3Q19", (16, 16, 64), 12, 17, walls, policy)
    whole, _, _ = native_run(g)
    parts = run_slabs(g, n_slabs, g["steps"], True)
    assert np.array_equal(parts, whole)


@pytest.mark.parametrize("n_slabs", [2, 4])
@pytest.mark.parametrize("split_faces", [False, True])
@pytest.mark.parametrize("name", ["cavity_d2q9_bgk_fp32", "cavity_d2q9_kbc_fp32", "channel2d_d2q9_bgk_outflow_fp32", "channel2d_d2q9_bgk_zouhe_pressure_fp32"])
def test_2d_slab_run_is_bit_identical(name, n_slabs, split_faces):
    """2-D x-slabs (reference: examples/cfd/lid_driven_cavity_2d_distributed.py): a D2Q9 call with a halo handle or a partial x range runs
    in the slab axis order (D2Q9X: physical x on the kernel's ghost-plane axis); it must give the bits of the undecomposed run, which
    uses the other axis order — lid-driven cavities (BGK, KBC) and channels with Regularized / Zou-He inlets and outflow / pressure
    outlets ON the slab faces."""
    from common import native_run

    g = load_golden(name)
    steps = 12
    whole, _, _ = native_run(g, steps=steps)
    parts = run_slabs(g, n_slabs, steps, split_faces)
    assert np.array_equal(parts, whole)


def test_a_dead_neighbour_is_a_loud_error():
    """A slab whose ring neighbour never signals: the device-side wait gives up after the time-out (never hanging the GPU), marks the handle in
    mapped host memory, and every later call on the handle — xlbn_step with the halo, push, signal, wait — fails with XLBN_E_STATE instead
    of stepping on stale ghost planes (round 1 let the run continue silently)."""
    g = load_golden("cavity_d3q19_bgk_fp32")
    stepper, f_0, f_1, bc_mask, missing_mask = native_case(g)
    handle = stepper._native_handle()
    L = native.lib()
    nx, ny, nz = g["shape"]
    st = native.stream_of(f_0)
    halos, bases = [], []
    for _ in range(2):
        h = C.c_void_p()
        native.check(L.xlbn_halo_create(stepper._lattice, stepper.precision_policy.store_precision.code, ny, nz, C.byref(h)))
        p, nbytes = C.c_void_p(), C.c_longlong()
        native.check(L.xlbn_halo_ghost_ptr(h, C.byref(p), C.byref(nbytes)))
        halos.append(h)
        bases.append(p.value)
    for i, h in enumerate(halos):
        other = C.create_string_buffer(np.uint64(bases[1 - i]).tobytes(), 64)
        native.check(L.xlbn_halo_connect(h, other, other, 1))
        native.check(L.xlbn_halo_set_timeout(h, 0.2))
    assert L.xlbn_halo_timed_out(halos[0]) == 0
    native.check(L.xlbn_halo_wait(halos[0], 7, st))  # nobody ever signals step 7
    torch.cuda.synchronize()
    assert L.xlbn_halo_timed_out(halos[0]) == 1 and L.xlbn_halo_timed_out(halos[1]) == 0
    dom = native.Domain(nx, ny, nz, 0, nx)
    rc = L.xlbn_step(handle, native.ptr(f_0), native.ptr(f_1), native.ptr(bc_mask), None, C.byref(dom), 1.0, 7, halos[0], st)
    assert rc == -5 and b"did not deliver" in L.xlbn_last_error()  # XLBN_E_STATE
    assert L.xlbn_halo_signal(halos[0], 8, st) == -5 and L.xlbn_halo_wait(halos[0], 8, st) == -5
    native.check(L.xlbn_step(handle, native.ptr(f_0), native.ptr(f_1), native.ptr(bc_mask), None, C.byref(dom), 1.0, 7, None, st))  # without the halo: fine
    torch.cuda.synchronize()
    for h in halos:
        L.xlbn_halo_destroy(h)
