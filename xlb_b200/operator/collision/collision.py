"""Base class of collision operators (reference: xlb/operator/collision/collision.py) + the shared native call."""

from xlb_b200 import native
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


def force3(force_vector, d):
    """Host double[3] from a length-d force vector (numpy / torch / sequence)."""
    import ctypes as C

    import numpy as np

    v = np.asarray(force_vector.detach().cpu().numpy() if hasattr(force_vector, "detach") else force_vector, dtype=np.float64).reshape(-1)
    if v.shape[0] != d:
        raise AssertionError("Check the dimensions of the input force!")  # forced_collision.py:30
    return (C.c_double * 3)(*(list(v) + [0.0] * (3 - d)))


class Collision(Operator):
    native_collision = None  # native.BGK / native.KBC / native.SMAGORINSKY_LES_BGK (| native.COLLISION_FORCED)
    native_smagorinsky = 0.17
    native_force = None  # ctypes double[3] for forced operators

    def _run_ext(self, f, feq, fout, rho, u, omega):
        """xlbn_collide_ext: every operator incl. SmagorinskyLESBGK and forced ones, on the GIVEN feq / rho / u."""
        vs = self.velocity_set
        for name, t in (("f", f), ("feq", feq), ("fout", fout)):
            native.require_cuda(t, name)
            if t.shape[0] != vs.q or t.shape != f.shape:
                raise ValueError(f"{type(self).__name__}: {name} has shape {tuple(t.shape)}, expected {tuple(f.shape)}")
        for name, t in (("rho", rho), ("u", u)):
            if t is not None:
                native.require_cuda(t, name)
        dims = native.dims_of(f, vs.d)
        native.check(
            native.lib().xlbn_collide_ext(
                self._lattice, self.native_collision, self._compute_code, native.ptr(f), native.dtype_code(f.dtype), native.ptr(feq),
                native.dtype_code(feq.dtype), native.ptr(fout), native.dtype_code(fout.dtype), native.ptr(rho),
                native.dtype_code(rho.dtype) if rho is not None else 0, native.ptr(u), native.dtype_code(u.dtype) if u is not None else 0,
                float(omega), self.native_force, float(self.native_smagorinsky), native.int3(dims), native.stream_of(f),
            )
        )  # fmt: skip
        return fout

    def _run(self, f, feq, fout, rho, omega):
        vs = self.velocity_set
        for name, t in (("f", f), ("feq", feq), ("fout", fout)):
            native.require_cuda(t, name)
            if t.shape[0] != vs.q or t.shape != f.shape:
                raise ValueError(f"{type(self).__name__}: {name} has shape {tuple(t.shape)}, expected {tuple(f.shape)}")
        if rho is not None:
            native.require_cuda(rho, "rho")
        dims = native.dims_of(f, vs.d)
        native.check(
            native.lib().xlbn_collide(
                self._lattice, self.native_collision, self._compute_code, native.ptr(f), native.dtype_code(f.dtype), native.ptr(feq),
                native.dtype_code(feq.dtype), native.ptr(fout), native.dtype_code(fout.dtype), native.ptr(rho),
                native.dtype_code(rho.dtype) if rho is not None else 0, float(omega), native.int3(dims), native.stream_of(f),
            )
        )  # fmt: skip
        return fout

    # reference signatures: JAX (f, feq, rho, u, omega) -> fout  (bgk.py:17-22, kbc.py:40-85)
    #                       Warp (f, feq, fout, rho, u, omega) -> fout (bgk.py:66-81, kbc.py:332-348)
    def _jax(self, f, feq, rho, u, omega):
        f = to_device_field(f)
        feq = to_device_field(feq, like=f)
        rho = to_device_field(rho, like=f) if rho is not None else None
        return self._run(f, feq, empty_like_field(f, self.velocity_set.q, f.dtype), rho, omega)

    def _warp(self, f, feq, fout, rho, u, omega):
        return self._run(f, feq, fout, rho, omega)
