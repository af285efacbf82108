"""Boundary maskers: index lists or a triangle mesh -> bc_mask / missing_mask (reference: xlb/operator/boundary_masker/)."""

from xlb_b200._exports import export

export(globals(), __name__, {"indices_boundary_masker": ["IndicesBoundaryMasker"], "mesh_boundary_masker": ["MeshBoundaryMasker"]})
