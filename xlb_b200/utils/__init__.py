"""I/O and visualisation helpers of the reference (xlb/utils/utils.py: PNG / VTK / USD writers, STL voxeliser) are
outside the scope of this backend.  The two names example scripts import are provided as
minimal, dependency-free writers so that those scripts run unchanged."""

from xlb_b200.utils.utils import save_image, save_fields_vtk
