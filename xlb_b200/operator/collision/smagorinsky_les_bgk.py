"""BGK collision with the Smagorinsky LES relaxation time; 3-D velocity sets.

Reference: xlb/operator/collision/smagorinsky_les_bgk.py — ctor L17-26, Warp functional L37-90, launch L124-138.  The
reference registers a WARP implementation only, with the argument order ``(f, feq, rho, u, fout, omega)`` (L124), and
its functional indexes ``c[2, l]`` (L71-76), so it exists for 3-D lattices only; both are kept.  The functional style
``(f, feq, rho, u, omega) -> fout`` is offered in addition under ComputeBackend.JAX.  Native: xlbn_collide_ext
(collide_smagorinsky in xlb_b200/csrc/lbm_math.cuh) and, inside the stepper, the fused kernel.
"""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision.collision import Collision
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class SmagorinskyLESBGK(Collision):
    native_collision = native.SMAGORINSKY_LES_BGK

    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None, smagorinsky_coef: float = 0.17):
        self.smagorinsky_coef = smagorinsky_coef
        super().__init__(velocity_set=velocity_set, precision_policy=precision_policy, compute_backend=compute_backend)
        if self.velocity_set.d != 3:
            raise NotImplementedError("SmagorinskyLESBGK: 3-D velocity sets only (the reference functional reads c[2, l])")

    @property
    def native_smagorinsky(self):
        return float(self.smagorinsky_coef)

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f, feq, rho, u, omega):
        f = to_device_field(f)
        feq = to_device_field(feq, like=f)
        return self._run_ext(f, feq, empty_like_field(f, self.velocity_set.q, f.dtype), None, None, omega)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, feq, rho, u, fout, omega):
        return self._run_ext(f, feq, fout, None, None, omega)
