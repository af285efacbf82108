"""Pi = sum_l c_l c_l f_l as a vector [xx, xy, xz, yy, yz, zz] (3-D) / [xx, xy, yy] (2-D).
Reference: xlb/operator/macroscopic/second_moment.py — JAX ``(fneq) -> pi`` L35-55, Warp ``(f, pi) -> pi`` L102-105."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.macroscopic._common import check_f
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class SecondMoment(Operator):
    def _run(self, f, pi):
        dims = check_f(self, f)
        native.require_cuda(pi, "pi")
        d = self.velocity_set.d
        if pi.shape[0] != d * (d + 1) // 2:
            raise ValueError(f"SecondMoment: pi must have {d * (d + 1) // 2} components")
        native.check(
            native.lib().xlbn_second_moment(
                self._lattice, self._compute_code, native.ptr(f), native.dtype_code(f.dtype), native.ptr(pi), native.dtype_code(pi.dtype),
                native.int3(dims), native.stream_of(f),
            )
        )  # fmt: skip
        return pi

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, fneq):
        fneq = to_device_field(fneq)
        d = self.velocity_set.d
        return self._run(fneq, empty_like_field(fneq, d * (d + 1) // 2, fneq.dtype))

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, pi):
        return self._run(f, pi)
