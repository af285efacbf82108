"""xlb_b200 — a B200-native (sm_100a) implementation of XLB's fused lattice-Boltzmann time step behind the
`xlb.operator` API.  `import xlb_b200 as xlb` (or plain `import xlb` through the alias package at the repo root)
gives the reference's import surface: enums, `init`, velocity sets, grid, operators, helpers, `distribute`.
All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/xlb_b200.h; there is no CPU path.
"""

import importlib

__version__ = "0.1.0"

from xlb_b200.compute_backend import ComputeBackend  # noqa: F401
from xlb_b200.default_config import DefaultConfig, init  # noqa: F401
from xlb_b200.grid_backend import GridBackend  # noqa: F401
from xlb_b200.physics_type import PhysicsType  # noqa: F401
from xlb_b200.precision_policy import Precision, PrecisionPolicy  # noqa: F401

# sub-packages that `import xlb` makes available as attributes (xlb.velocity_set.D3Q19, xlb.operator.stepper, ...)
for _name in (
    "velocity_set", "grid", "helper", "utils", "distribute",
    "operator.equilibrium", "operator.collision", "operator.stream", "operator.macroscopic", "operator.force",
    "operator.boundary_condition", "operator.boundary_masker", "operator.stepper",
):  # fmt: skip
    importlib.import_module(f"{__name__}.{_name}")
del _name
