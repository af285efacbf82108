"""Builds ``bc_mask`` (uint8 id per cell) and ``missing_mask`` (bool per direction per cell) from per-BC index lists.

Reference: xlb/operator/boundary_masker/indices_boundary_masker.py — JAX L45-101, Warp L103-224.  The two reference
backends use different algorithms that agree everywhere except in ``missing_mask`` on BC-free domain-face cells
(SURVEY.md §8a row M1).  Both are reproduced bit-exactly on the device (xlb_b200/csrc/masker.cu); the operator's
``compute_backend`` selects which one, so results match whichever reference backend a script was written for.

Call: ``masker(bclist, bc_mask, missing_mask, start_index=None) -> (bc_mask, missing_mask)``.
Slab decomposition: ``start_index`` is the global coordinate of local cell 0 and ``global_shape`` (extra keyword) the
extents of the whole domain; indices are always global.
"""

import numpy as np
import torch

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.operator import Operator
from xlb_b200.operator.stream.stream import Stream


class IndicesBoundaryMasker(Operator):
    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None):
        super().__init__(velocity_set, precision_policy, compute_backend)
        self.stream = Stream(self.velocity_set, self.precision_policy, self.compute_backend)

    def are_indices_in_interior(self, indices, shape):
        """Per index: strictly inside the domain, not on its boundary (reference L29-40)."""
        d = self.velocity_set.d
        shape_array = np.array(shape)
        indices = np.asarray(indices)
        return np.all((indices[:d] > 0) & (indices[:d] < shape_array[:d, np.newaxis] - 1), axis=0)

    def _run(self, bclist, bc_mask, missing_mask, start_index, global_shape, mode):
        vs = self.velocity_set
        native.require_cuda(bc_mask, "bc_mask")
        native.require_cuda(missing_mask, "missing_mask")
        if bc_mask.dtype != torch.uint8 or missing_mask.dtype != torch.bool:
            raise TypeError("bc_mask must be uint8 and missing_mask bool")
        local = native.dims_of(missing_mask, vs.d)
        if missing_mask.shape[0] != vs.q or bc_mask.shape[0] != 1 or native.dims_of(bc_mask, vs.d) != local:
            raise ValueError("bc_mask / missing_mask shapes do not match the velocity set")
        start = tuple(start_index) + (0,) * (3 - len(start_index)) if start_index is not None else (0, 0, 0)
        gshape = tuple(global_shape) + (1,) * (3 - len(global_shape)) if global_shape is not None else local
        stream = native.stream_of(bc_mask)
        L = native.lib()

        solid = incoming = None
        if mode == native.MASK_JAX:
            if bool(missing_mask.any()):  # entries set by an earlier masker are streamed along, as the reference does (L56-63, 92)
                incoming = missing_mask.clone()
            solid = torch.zeros((local[0] + 2) * (local[1] + 2) * (local[2] + 2), dtype=torch.uint8, device=bc_mask.device)
        for bc in bclist:
            assert bc.indices is not None, f'Please specify indices associated with the {bc.__class__.__name__} BC using keyword "indices"!'
            assert bc.mesh_vertices is None, f"Please use MeshBoundaryMasker operator if {bc.__class__.__name__} is imposed on a mesh (e.g. STL)!"
            idx = np.asarray(bc.indices, dtype=np.int64)
            if idx.ndim != 2 or idx.shape[0] != vs.d:
                raise ValueError(f"{bc.__class__.__name__}: indices must have shape ({vs.d}, n), got {idx.shape}")
            n = idx.shape[1]
            if n == 0:
                continue
            if idx.shape[0] == 2:
                idx = np.vstack([idx, np.zeros((1, n), dtype=np.int64)])
            flag = bool(bc.needs_padding)
            if mode == native.MASK_JAX:  # the JAX masker pads when ANY index of the BC is interior (reference L75)
                flag = flag and bool(np.any(self.are_indices_in_interior(idx, gshape)))
            d_idx = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32), device=bc_mask.device)
            native.check(
                L.xlbn_mask_indices(
                    self._lattice, mode, native.ptr(d_idx), n, int(bc.id), int(flag), native.int3(gshape), native.int3(start), native.int3(local),
                    native.ptr(bc_mask), native.ptr(missing_mask), native.ptr(solid), stream,
                )
            )  # fmt: skip
        if mode == native.MASK_JAX:
            native.check(
                L.xlbn_mask_finalize_jax(self._lattice, native.int3(gshape), native.int3(start), native.int3(local), native.ptr(missing_mask), native.ptr(solid),
                                         native.ptr(incoming), stream)
            )
        return bc_mask, missing_mask

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, bclist, bc_mask, missing_mask, start_index=None, global_shape=None):
        return self._run(bclist, bc_mask, missing_mask, start_index, global_shape, native.MASK_JAX)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, bclist, bc_mask, missing_mask, start_index=None, global_shape=None):
        return self._run(bclist, bc_mask, missing_mask, start_index, global_shape, native.MASK_WARP)
