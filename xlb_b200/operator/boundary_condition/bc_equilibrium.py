"""EquilibriumBC: all populations of the BC cells are set to feq(rho, u) (streaming step).
Reference: xlb/operator/boundary_condition/bc_equilibrium.py:24-101."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.boundary_condition.boundary_condition import BoundaryCondition, ImplementationStep
from xlb_b200.operator.equilibrium import Equilibrium, QuadraticEquilibrium
from xlb_b200.operator.operator import Operator


class EquilibriumBC(BoundaryCondition):
    native_kind = native.BC_EQUILIBRIUM

    def __init__(self, rho, u, equilibrium_operator=None, velocity_set=None, precision_policy=None, compute_backend=None, indices=None, mesh_vertices=None):
        self.rho = rho
        self.u = u
        self.equilibrium_operator = QuadraticEquilibrium() if equilibrium_operator is None else equilibrium_operator
        if not issubclass(type(self.equilibrium_operator), Equilibrium):
            raise ValueError("Equilibrium operator must be a subclass of Equilibrium")
        super().__init__(ImplementationStep.STREAMING, velocity_set, precision_policy, compute_backend, indices, mesh_vertices)
        if len(self.u) < self.velocity_set.d:
            raise ValueError(f"EquilibriumBC: u needs {self.velocity_set.d} components")

    def native_desc(self):
        d = super().native_desc()
        d.rho = float(self.rho)
        for a in range(self.velocity_set.d):
            d.u[a] = float(self.u[a])
        return d

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_jax(f_pre, f_post, bc_mask, missing_mask)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_warp(f_pre, f_post, bc_mask, missing_mask)
