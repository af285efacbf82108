"""Extrapolation outflow (Geier et al. 2015): missing populations are taken from auxiliary values prepared after the
previous collision, f_out[opp[l]] = (1 - 1/sqrt3) f_post_stream[l] + 1/sqrt3 * (neighbour's post-stream f_l), stored in
the populations that leave the domain.
Reference: xlb/operator/boundary_condition/bc_extrapolation_outflow.py — normal from index statistics L63-77,
streaming part L120-129 / L152-170, post-collision aux update L91-118 / L172-195 (runs inside the fused step kernel:
xlb_b200/csrc/step_kernel.cuh, bc_cell)."""

from collections import Counter

import numpy as np

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.boundary_condition.boundary_condition import BoundaryCondition, ImplementationStep
from xlb_b200.operator.operator import Operator


class ExtrapolationOutflowBC(BoundaryCondition):
    native_kind = native.BC_EXTRAPOLATION_OUTFLOW

    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None, indices=None, mesh_vertices=None):
        super().__init__(ImplementationStep.STREAMING, velocity_set, precision_policy, compute_backend, indices, mesh_vertices)
        self.normal = None
        if indices is not None:
            self._get_normal_vec(indices)

    def _get_normal_vec(self, indices):
        # most common coordinate per axis; the axis on which (almost) all cells agree is the face normal
        freq = [Counter(np.asarray(coord).tolist()).most_common(1)[0] for coord in indices]
        counts = np.array([count for _, count in freq])
        elements = np.array([element for element, _ in freq])
        self.normal = counts // counts.max()
        if elements[np.argmax(counts)] == 0:
            self.normal = self.normal * -1

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_jax(f_pre, f_post, bc_mask, missing_mask)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_warp(f_pre, f_post, bc_mask, missing_mask)
