"""Second-order Hermite equilibrium  feq_l = rho w_l (1 + cu (1 + cu/2) - 1.5 u.u),  cu = 3 c_l.u.

Reference: xlb/operator/equilibrium/quadratic_equilibrium.py — JAX ``(rho, u) -> feq`` L18-25 (result in the dtype of
the inputs), Warp ``(rho, u, f) -> f`` L85-97.  Native: xlbn_equilibrium (arithmetic in the policy's compute dtype).
"""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.equilibrium.equilibrium import Equilibrium
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class QuadraticEquilibrium(Equilibrium):
    def _run(self, rho, u, f):
        for name, t in (("rho", rho), ("u", u), ("f", f)):
            native.require_cuda(t, name)
        vs = self.velocity_set
        if rho.shape[0] != 1 or u.shape[0] != vs.d or f.shape[0] != vs.q or rho.shape[1:] != u.shape[1:] or rho.shape[1:] != f.shape[1:]:
            raise ValueError(f"QuadraticEquilibrium: shapes rho {tuple(rho.shape)}, u {tuple(u.shape)}, f {tuple(f.shape)} do not match {vs}")
        dims = native.dims_of(f, vs.d)
        native.check(
            native.lib().xlbn_equilibrium(
                self._lattice, self._compute_code, native.ptr(rho), native.dtype_code(rho.dtype), native.ptr(u), native.dtype_code(u.dtype),
                native.ptr(f), native.dtype_code(f.dtype), native.int3(dims), native.stream_of(f),
            )
        )  # fmt: skip
        return f

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, rho, u):
        u = to_device_field(u)
        rho = to_device_field(rho, like=u)
        return self._run(rho, u, empty_like_field(u, self.velocity_set.q, u.dtype))

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, rho, u, f):
        return self._run(rho, u, f)
