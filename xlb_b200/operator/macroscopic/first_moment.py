"""u = (sum_l c_l f_l) / rho with a caller-provided density.
Reference: xlb/operator/macroscopic/first_moment.py — JAX ``(f, rho) -> u`` L14-18, Warp ``(f, rho, u) -> u`` L61-67."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.macroscopic._common import check_f
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class FirstMoment(Operator):
    def _run(self, f, rho, u):
        dims = check_f(self, f)
        native.require_cuda(rho, "rho")
        native.require_cuda(u, "u")
        native.check(
            native.lib().xlbn_first_moment(
                self._lattice, self._compute_code, native.ptr(f), native.dtype_code(f.dtype), native.ptr(rho), native.dtype_code(rho.dtype),
                native.ptr(u), native.dtype_code(u.dtype), native.int3(dims), native.stream_of(f),
            )
        )  # fmt: skip
        return u

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f, rho):
        f = to_device_field(f)
        rho = to_device_field(rho, like=f)
        return self._run(f, rho, empty_like_field(f, self.velocity_set.d, f.dtype))

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, rho, u):
        return self._run(f, rho, u)
