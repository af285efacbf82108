"""A collision operator followed by a body-force term.

Reference: xlb/operator/collision/forced_collision.py — ctor L16-33, JAX L35-39 (``fout = collision(...); fout =
forcing(fout, feq, rho, u)``), Warp functional L46-50, launch L89-103.  Only the "exact_difference" scheme exists
(L27).  Native: xlbn_collide_ext with XLBN_COLLISION_FORCED or-ed onto the wrapped operator's code — one kernel for
both stages — and, inside the stepper, the fused kernel (xlbn_stepper_set_force).
"""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision.collision import Collision, force3
from xlb_b200.operator.force.exact_difference_force import ExactDifference
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class ForcedCollision(Collision):
    def __init__(self, collision_operator: Operator, forcing_scheme="exact_difference", force_vector=None):
        assert collision_operator is not None
        self.collision_operator = collision_operator
        super().__init__(
            velocity_set=collision_operator.velocity_set,
            precision_policy=collision_operator.precision_policy,
            compute_backend=collision_operator.compute_backend,
        )
        assert forcing_scheme == "exact_difference", NotImplementedError(f"Force model {forcing_scheme} not implemented!")
        assert force_vector is not None and force_vector.shape[0] == self.velocity_set.d, "Check the dimensions of the input force!"
        self.force_vector = force_vector
        self.forcing_operator = ExactDifference(
            force_vector, velocity_set=self.velocity_set, precision_policy=self.precision_policy, compute_backend=self.compute_backend
        )
        self.native_force = force3(force_vector, self.velocity_set.d)

    @property
    def native_collision(self):
        return self.collision_operator.native_collision | native.COLLISION_FORCED

    @property
    def native_smagorinsky(self):
        return self.collision_operator.native_smagorinsky

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f, feq, rho, u, omega):
        f = to_device_field(f)
        feq, rho, u = to_device_field(feq, like=f), to_device_field(rho, like=f), to_device_field(u, like=f)
        return self._run_ext(f, feq, empty_like_field(f, self.velocity_set.q, f.dtype), rho, u, omega)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f, feq, fout, rho, u, omega):
        return self._run_ext(f, feq, fout, rho, u, omega)
