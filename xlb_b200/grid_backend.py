"""Grid-backend enum kept for import compatibility (reference: xlb/grid_backend.py)."""

from enum import Enum, auto


class GridBackend(Enum):
    JAX = auto()
    WARP = auto()
    OOC = auto()
