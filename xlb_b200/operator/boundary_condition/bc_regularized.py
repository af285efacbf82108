"""Regularized boundary condition: Zou-He followed by a regularisation of ALL populations from the non-equilibrium
momentum flux, f_l = feq_l + 4.5 w_l (Q_l : Pi_neq)  (Latt et al. 2008).
Reference: xlb/operator/boundary_condition/bc_regularized.py — ctor L43-65, JAX L67-124, Warp L134-202."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.boundary_condition.bc_zouhe import ZouHeBC
from xlb_b200.operator.macroscopic import SecondMoment as MomentumFlux
from xlb_b200.operator.operator import Operator


class RegularizedBC(ZouHeBC):
    _kind_by_type = {"velocity": native.BC_REGULARIZED_VELOCITY, "pressure": native.BC_REGULARIZED_PRESSURE}

    def __init__(self, bc_type, profile=None, prescribed_value=None, velocity_set=None, precision_policy=None, compute_backend=None, indices=None, mesh_vertices=None):
        super().__init__(bc_type, profile, prescribed_value, velocity_set, precision_policy, compute_backend, indices, mesh_vertices)
        self.momentum_flux = MomentumFlux(self.velocity_set, self.precision_policy, self.compute_backend)

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_jax(f_pre, f_post, bc_mask, missing_mask)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_warp(f_pre, f_post, bc_mask, missing_mask)
