"""FullwayBounceBackBC: after collision every population of a BC cell is replaced by the post-streaming population of the opposite direction.
Reference: xlb/operator/boundary_condition/bc_fullway_bounce_back.py:22-86."""

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.boundary_condition.boundary_condition import BoundaryCondition, ImplementationStep
from xlb_b200.operator.operator import Operator


class FullwayBounceBackBC(BoundaryCondition):
    native_kind = native.BC_FULLWAY_BOUNCE_BACK

    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None, indices=None, mesh_vertices=None):
        super().__init__(ImplementationStep.COLLISION, velocity_set, precision_policy, compute_backend, indices, mesh_vertices)
        # interior geometries need their neighbours tagged to find the missing directions
        self.needs_padding = False

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_jax(f_pre, f_post, bc_mask, missing_mask)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_warp(f_pre, f_post, bc_mask, missing_mask)
