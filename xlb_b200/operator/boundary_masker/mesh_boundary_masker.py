"""Masks for a boundary condition that is given as a triangle mesh (e.g. an STL body in a wind tunnel); 3-D only.

Reference: xlb/operator/boundary_masker/mesh_boundary_masker.py — ctor L15-32, Warp kernel L153-190 (voxels overlapped
by a triangle become solid, id 255; every other cell with a solid neighbour in direction l becomes a boundary cell with
``missing_mask[opp[l]] = True``), host side L193-236 (argument checks, mesh inside the domain, one flat vertex list with
three consecutive rows per triangle).  The reference finds triangle / voxel pairs with Warp's BVH (``wp.Mesh``,
``wp.mesh_query_aabb``); here one warp per triangle walks the voxels under the triangle's bounding box
(xlb_b200/csrc/mesh_masker.cu, ``xlbn_mask_mesh``) — same pairs, no tree.

``edge_test`` selects the triangle / unit-box overlap test:

* ``"schwarz_seidel"`` (default): the published test the reference cites (Schwarz & Seidel 2010);
* ``"reference"``: the reference's ``pre_compute`` exactly as written (L78-97).  There both components of every edge
  normal are read from ``edges[i][axis0]`` and both offsets from ``verts[i][axis0]``, so the edge functions no longer
  depend on the triangle and only voxels with ``|i-j|, |j-k|, |k-i| <= 1`` can pass.  This mode reproduces the reference's
  masks bit for bit (tests/golden/warp_mesh_*.npz) and exists for that comparison; it does not voxelise a body.

Call: ``masker(bc, bc_mask, missing_mask) -> (bc_mask, missing_mask)`` for ONE mesh-based BC (reference L193-198).
Slab decomposition: ``start_index`` (global coordinate of local cell 0) and ``global_shape`` as extra keywords; the mesh is
given in GLOBAL grid units, every rank voxelises the triangles that reach its slab (plus one cell of halo, so that boundary
cells next to a solid voxel of the neighbouring slab are found).
"""

import numpy as np
import torch

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.operator import Operator

EDGE_TESTS = {"schwarz_seidel": 0, "reference": 1}


class MeshBoundaryMasker(Operator):
    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None, edge_test="schwarz_seidel"):
        super().__init__(velocity_set, precision_policy, compute_backend)
        if self.velocity_set.d == 2:
            raise NotImplementedError("This Operator is not implemented in 2D!")
        if edge_test not in EDGE_TESTS:
            raise ValueError(f"edge_test must be one of {sorted(EDGE_TESTS)}, got {edge_test!r}")
        self.edge_test = edge_test

    def _run(self, bc, bc_mask, missing_mask, start_index=None, global_shape=None):
        vs = self.velocity_set
        assert bc.mesh_vertices is not None, f'Please provide the mesh vertices for {bc.__class__.__name__} BC using keyword "mesh_vertices"!'
        assert bc.indices is None, f"Please use IndicesBoundaryMasker operator if {bc.__class__.__name__} is imposed on known indices of the grid!"
        mesh_vertices = np.asarray(bc.mesh_vertices)
        assert mesh_vertices.ndim == 2 and mesh_vertices.shape[1] == vs.d, "Mesh points must be reshaped into an array (N, 3) where N indicates number of points!"
        if mesh_vertices.shape[0] % 3:
            raise ValueError(f"mesh_vertices has {mesh_vertices.shape[0]} rows: three consecutive rows per triangle are expected")
        if bc_mask.dtype != torch.uint8 or missing_mask.dtype != torch.bool:
            raise TypeError("bc_mask must be uint8 and missing_mask bool")
        dims = native.dims_of(missing_mask, vs.d)
        if missing_mask.shape[0] != vs.q or bc_mask.shape[0] != 1 or native.dims_of(bc_mask, vs.d) != dims:
            raise ValueError("bc_mask / missing_mask shapes do not match the velocity set")
        domain = tuple(global_shape) if global_shape is not None else tuple(dims)
        mesh_min, mesh_max = mesh_vertices.min(axis=0), mesh_vertices.max(axis=0)
        if any(mesh_min < 0) or any(mesh_max >= np.array(domain)):
            raise ValueError(
                f"Mesh extents ({mesh_min}, {mesh_max}) exceed domain dimensions {domain}. The mesh must be fully contained within the domain."
            )
        if start_index is not None:  # local coordinates of this slab; triangles elsewhere fall outside the padded volume and are skipped
            mesh_vertices = mesh_vertices - np.asarray(start_index, dtype=mesh_vertices.dtype)[: vs.d]
        assert not getattr(bc, "needs_mesh_distance", False), 'Please use "MeshDistanceBoundaryMasker" if this BC needs mesh distance!'
        native.require_cuda(bc_mask, "bc_mask")
        native.require_cuda(missing_mask, "missing_mask")
        bc.__dict__.pop("mesh_vertices", None)  # reference L212-213: the BC is done with its vertices

        verts = torch.as_tensor(np.ascontiguousarray(mesh_vertices, dtype=np.float32), device=bc_mask.device)
        solid = torch.empty((dims[0] + 2) * (dims[1] + 2) * (dims[2] + 2), dtype=torch.uint8, device=bc_mask.device)
        native.check(
            native.lib().xlbn_mask_mesh(
                self._lattice, native.ptr(verts), mesh_vertices.shape[0] // 3, int(bc.id), EDGE_TESTS[self.edge_test], native.int3(dims),
                native.ptr(bc_mask), native.ptr(missing_mask), native.ptr(solid), native.stream_of(bc_mask),
            )
        )  # fmt: skip
        return bc_mask, missing_mask

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, bc, bc_mask, missing_mask, start_index=None, global_shape=None):
        return self._run(bc, bc_mask, missing_mask, start_index, global_shape)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, bc, bc_mask, missing_mask, start_index=None, global_shape=None):
        return self._run(bc, bc_mask, missing_mask, start_index, global_shape)
