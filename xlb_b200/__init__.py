"""xlb_b200 — a B200-native (sm_100a) implementation of XLB's fused lattice-Boltzmann time step behind the
`xlb.operator` API.  `import xlb_b200 as xlb` (or plain `import xlb` through the alias package at the repo root)
gives the reference's import surface: enums, `init`, velocity sets, grid, operators, helpers, `distribute`.
All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/xlb_b200.h; there is no CPU path.
"""

__version__ = "0.1.0"

# Enum classes
from xlb_b200.compute_backend import ComputeBackend as ComputeBackend
from xlb_b200.precision_policy import PrecisionPolicy as PrecisionPolicy, Precision as Precision
from xlb_b200.physics_type import PhysicsType as PhysicsType
from xlb_b200.grid_backend import GridBackend as GridBackend

# Config
from xlb_b200.default_config import init as init, DefaultConfig as DefaultConfig

# Velocity sets, operators, grid, helpers, utils, distribution
import xlb_b200.velocity_set
import xlb_b200.operator.equilibrium
import xlb_b200.operator.collision
import xlb_b200.operator.stream
import xlb_b200.operator.boundary_condition
import xlb_b200.operator.boundary_masker
import xlb_b200.operator.macroscopic
import xlb_b200.operator.stepper
import xlb_b200.operator.force
import xlb_b200.grid
import xlb_b200.helper
import xlb_b200.utils
import xlb_b200.distribute
