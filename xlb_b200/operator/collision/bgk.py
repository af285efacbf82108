"""BGK collision  fout = f - omega (f - feq).  Reference: xlb/operator/collision/bgk.py:12-81.  Native: xlbn_collide."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision.collision import Collision
from xlb_b200.operator.operator import Operator


class BGK(Collision):
    native_collision = native.BGK

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f, feq, rho, u, omega):
        return self._jax(f, feq, None, u, omega)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, feq, fout, rho, u, omega):
        return self._warp(f, feq, fout, None, u, omega)
