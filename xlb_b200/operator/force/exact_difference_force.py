"""Exact-difference body force (Kupershtokh 2004): f += feq(rho, u + F) - feq(rho, u).

Reference: xlb/operator/force/exact_difference_force.py — ctor L26-43, JAX L45-70, Warp functional L79-84, launch
L117-125.  A single constant force vector (L33: no spatially varying field).  Native: xlbn_exact_difference
(exact_difference in xlb_b200/csrc/lbm_math.cuh).
"""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import empty_like_field, to_device_field


class ExactDifference(Operator):
    def __init__(self, force_vector, equilibrium: Operator = None, velocity_set=None, precision_policy=None, compute_backend=None):
        from xlb_b200.operator.collision.collision import force3
        from xlb_b200.operator.equilibrium import QuadraticEquilibrium

        self.force_vector = force_vector
        super().__init__(velocity_set, precision_policy, compute_backend)
        self.equilibrium = QuadraticEquilibrium(self.velocity_set, self.precision_policy, self.compute_backend) if equilibrium is None else equilibrium
        self._force3 = force3(force_vector, self.velocity_set.d)

    def _run(self, f_postcollision, feq, fout, rho, u):
        vs = self.velocity_set
        for name, t in (("f_postcollision", f_postcollision), ("feq", feq), ("fout", fout), ("rho", rho), ("u", u)):
            native.require_cuda(t, name)
        if feq.shape != f_postcollision.shape or fout.shape != f_postcollision.shape or f_postcollision.shape[0] != vs.q:
            raise ValueError(f"ExactDifference: populations must all be [q={vs.q}, ...] of one shape")
        dims = native.dims_of(f_postcollision, vs.d)
        native.check(
            native.lib().xlbn_exact_difference(
                self._lattice, self._compute_code, native.ptr(f_postcollision), native.dtype_code(f_postcollision.dtype), native.ptr(feq),
                native.dtype_code(feq.dtype), native.ptr(fout), native.dtype_code(fout.dtype), native.ptr(rho), native.dtype_code(rho.dtype),
                native.ptr(u), native.dtype_code(u.dtype), self._force3, native.int3(dims), native.stream_of(f_postcollision),
            )
        )  # fmt: skip
        return fout

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_postcollision, feq, rho, u):
        f = to_device_field(f_postcollision)
        feq, rho, u = to_device_field(feq, like=f), to_device_field(rho, like=f), to_device_field(u, like=f)
        return self._run(f, feq, empty_like_field(f, self.velocity_set.q, f.dtype), rho, u)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_postcollision, feq, fout, rho, u):
        return self._run(f_postcollision, feq, fout, rho, u)
