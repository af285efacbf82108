"""Boundary maskers (index lists -> bc_mask / missing_mask).  The mesh masker of the reference is out of scope."""

from xlb_b200._exports import export

export(globals(), __name__, {"indices_boundary_masker": ["IndicesBoundaryMasker"]})
