"""I/O and visualisation helpers of the reference (xlb/utils/utils.py: PNG / VTK / USD writers, STL voxeliser) are
outside the scope of this backend.  The two names example scripts import are provided as
minimal, dependency-free writers so that those scripts run unchanged, plus an STL reader for mesh-based BCs (the
reference scripts use `trimesh`, which this image does not have)."""

from xlb_b200.utils.utils import read_stl, save_fields_vtk, save_image
