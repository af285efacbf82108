"""Kept so that `from xlb.grid_backend import GridBackend` keeps working; the framework has a single grid type."""

from enum import Enum

GridBackend = Enum("GridBackend", ["JAX", "WARP", "OOC"])
