"""rho = sum_l f_l.  Reference: xlb/operator/macroscopic/zero_moment.py — JAX ``(f) -> rho`` L14-17, Warp ``(f, rho) -> rho`` L47-49."""

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.macroscopic._common import run_macroscopic
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class ZeroMoment(Operator):
    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f):
        f = to_device_field(f)
        rho = empty_like_field(f, 1, f.dtype)
        run_macroscopic(self, f, rho, None)
        return rho

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, rho):
        run_macroscopic(self, f, rho, None)
        return rho
