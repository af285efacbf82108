"""KBC (Karlin-Boesch-Chikatamarla) entropic multi-relaxation collision; D3Q27 and D2Q9 only.

Reference: xlb/operator/collision/kbc.py — ctor L25-38, JAX L40-85, Warp functional L268-296; unsupported lattices
raise NotImplementedError (L71-72, L184-185).  Native: xlbn_collide (collide_kbc in xlb_b200/csrc/lbm_math.cuh).
"""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision.collision import Collision
from xlb_b200.operator.macroscopic import SecondMoment as MomentumFlux
from xlb_b200.operator.operator import Operator
from xlb_b200.velocity_set import D2Q9, D3Q27


class KBC(Collision):
    native_collision = native.KBC

    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None):
        super().__init__(velocity_set=velocity_set, precision_policy=precision_policy, compute_backend=compute_backend)
        if not isinstance(self.velocity_set, (D3Q27, D2Q9)):
            raise NotImplementedError("Velocity set not supported: {}".format(type(self.velocity_set)))
        self.momentum_flux = MomentumFlux(self.velocity_set, self.precision_policy, self.compute_backend)
        self.epsilon = 1e-32

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f, feq, rho, u, omega):
        return self._jax(f, feq, rho, u, omega)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, feq, fout, rho, u, omega):
        return self._warp(f, feq, fout, rho, u, omega)
