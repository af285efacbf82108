"""Base class of boundary conditions.

Reference: xlb/operator/boundary_condition/boundary_condition.py — ctor, id registration and capability flags L28-73,
generic stand-alone kernel L83-117, aux-data initialisation L119-175.

Every BC is described to the native library by ``(id, kind, rho, u)`` (xlbn_bc_desc, include/xlb_b200.h).  Inside the
fused step the BC is selected per cell through the uint8 ``bc_mask``; stand-alone ``bc(f_pre, f_post, bc_mask,
missing_mask) -> f_post`` runs xlbn_bc_apply, the counterpart of the reference's generic BC kernel.
"""

from enum import Enum, auto

import numpy as np
import torch

from xlb_b200 import native
from xlb_b200.default_config import DefaultConfig
from xlb_b200.operator.boundary_condition.boundary_condition_registry import boundary_condition_registry
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import to_device_field


class ImplementationStep(Enum):
    COLLISION = auto()
    STREAMING = auto()


class BoundaryCondition(Operator):
    native_kind = native.BC_NONE

    def __init__(self, implementation_step, velocity_set=None, precision_policy=None, compute_backend=None, indices=None, mesh_vertices=None):
        self.id = boundary_condition_registry.register_boundary_condition(self.__class__.__name__ + "_" + str(hash(self)))
        velocity_set = velocity_set or DefaultConfig.velocity_set
        precision_policy = precision_policy or DefaultConfig.default_precision_policy
        compute_backend = compute_backend or DefaultConfig.default_backend
        super().__init__(velocity_set, precision_policy, compute_backend)

        self.indices = indices
        self.mesh_vertices = mesh_vertices
        self.implementation_step = implementation_step
        # capability flags, same meaning as the reference (boundary_condition.py:54-73)
        self.needs_padding = False
        self.needs_mesh_distance = False
        self.needs_aux_init = False
        self.is_initialized_with_aux_data = False
        self.num_of_aux_data = 0
        self.needs_aux_recovery = False

    # -- native description -----------------------------------------------------------------------------------
    def native_desc(self) -> native.BcDesc:
        d = native.BcDesc()
        d.id, d.kind, d.rho = int(self.id), int(self.native_kind), 1.0
        d.u[0] = d.u[1] = d.u[2] = 0.0
        return d

    def update_bc_auxilary_data(self, f_pre, f_post, bc_mask, missing_mask):
        """Post-collision aux hook (reference: boundary_condition.py:75-81); a no-op except for ExtrapolationOutflowBC,
        whose update runs inside the fused step kernel."""
        return f_post

    # -- stand-alone application: (f_pre, f_post, bc_mask, missing_mask) -> f_post --------------------------------
    def _apply(self, f_pre, f_post, bc_mask, missing_mask):
        vs = self.velocity_set
        for name, t in (("f_pre", f_pre), ("f_post", f_post), ("bc_mask", bc_mask), ("missing_mask", missing_mask)):
            native.require_cuda(t, name)
        if f_pre.shape != f_post.shape or f_pre.dtype != f_post.dtype or f_pre.shape[0] != vs.q:
            raise ValueError(f"{type(self).__name__}: f_pre / f_post shapes {tuple(f_pre.shape)} / {tuple(f_post.shape)}")
        if bc_mask.dtype != torch.uint8 or missing_mask.dtype != torch.bool:
            raise TypeError("bc_mask must be uint8 and missing_mask bool")
        if bc_mask.shape[1:] != f_pre.shape[1:] or missing_mask.shape != (vs.q,) + tuple(f_pre.shape[1:]):
            raise ValueError("bc_mask / missing_mask do not match the field shape")
        dims = native.dims_of(f_pre, vs.d)
        desc = self.native_desc()
        native.check(
            native.lib().xlbn_bc_apply(
                self._lattice, self._compute_code, desc, native.ptr(f_pre), native.ptr(f_post), native.dtype_code(f_pre.dtype),
                native.ptr(bc_mask), native.ptr(missing_mask), native.int3(dims), native.stream_of(f_pre),
            )
        )  # fmt: skip
        return f_post

    def _call_jax(self, f_pre, f_post, bc_mask, missing_mask):
        f_post = to_device_field(f_post)
        f_pre = to_device_field(f_pre, like=f_post)
        return self._apply(f_pre, f_post.copy(), to_device_field(bc_mask, like=f_post), to_device_field(missing_mask, like=f_post))

    def _call_warp(self, f_pre, f_post, bc_mask, missing_mask):
        return self._apply(f_pre, f_post, bc_mask, missing_mask)

    # -- aux data (prescribed values) ---------------------------------------------------------------------------
    def _prescribed_values_at(self, cells_global: np.ndarray, missing_cells: np.ndarray, global_shape) -> np.ndarray:
        raise NotImplementedError

    def aux_data_init(self, f_0, f_1, bc_mask, missing_mask, start_index=None, global_shape=None):
        """Encode the prescribed scalar of every cell of this BC into ``f_1[0, cell]`` in the STORE dtype
        (reference: boundary_condition.py:119-175; only the first aux value is used by the in-scope BCs)."""
        native.require_cuda(f_1, "f_1")
        cells = torch.nonzero(bc_mask[0] == self.id, as_tuple=False)  # (n, 2|3) local coordinates
        if cells.shape[0] > 0:
            idx = tuple(cells[:, a] for a in range(cells.shape[1]))
            missing_cells = missing_mask[(slice(None),) + idx].cpu().numpy()  # (q, n)
            local = cells.t().cpu().numpy().astype(np.int64)
            if start_index is not None:
                local[: len(start_index)] += np.asarray(start_index, dtype=np.int64)[:, None]
            if global_shape is None:
                global_shape = tuple(f_1.shape[1 : 1 + self.velocity_set.d])
            values = np.asarray(self._prescribed_values_at(local, missing_cells, tuple(global_shape)), dtype=np.float64).reshape(-1)
            f_1[(0,) + idx] = torch.as_tensor(values, device=f_1.device).to(f_1.dtype)
        self.is_initialized_with_aux_data = True
        return f_0, f_1
