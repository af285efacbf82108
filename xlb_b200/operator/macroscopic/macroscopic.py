"""(rho, u) from the populations.
Reference: xlb/operator/macroscopic/macroscopic.py — JAX ``(f) -> (rho, u)`` L26-31, Warp ``(f, rho, u) -> (rho, u)`` L58-65."""

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.macroscopic._common import run_macroscopic
from xlb_b200.operator.macroscopic.first_moment import FirstMoment
from xlb_b200.operator.macroscopic.zero_moment import ZeroMoment
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class Macroscopic(Operator):
    def __init__(self, *args, **kwargs):
        self.zero_moment = ZeroMoment(*args, **kwargs)
        self.first_moment = FirstMoment(*args, **kwargs)
        super().__init__(*args, **kwargs)

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f):
        f = to_device_field(f)
        rho = empty_like_field(f, 1, f.dtype)
        u = empty_like_field(f, self.velocity_set.d, f.dtype)
        run_macroscopic(self, f, rho, u)
        return rho, u

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, rho, u):
        run_macroscopic(self, f, rho, u)
        return rho, u
