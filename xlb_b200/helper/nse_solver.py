"""Field allocation for the Navier-Stokes stepper (reference: xlb/helper/nse_solver.py:7-41)."""

from typing import Tuple

from xlb_b200.default_config import DefaultConfig
from xlb_b200.grid import grid_factory
from xlb_b200.precision_policy import Precision


def create_nse_fields(grid_shape: Tuple[int, int, int] = None, grid=None, velocity_set=None, compute_backend=None, precision_policy=None):
    """Returns (grid, f_0, f_1, missing_mask, bc_mask): two population buffers in the store dtype, the bool
    missing-direction mask [q, ...] and the uint8 boundary-id mask [1, ...]."""
    velocity_set = velocity_set or DefaultConfig.velocity_set
    precision_policy = precision_policy or DefaultConfig.default_precision_policy
    if grid is None:
        if grid_shape is None:
            raise ValueError("grid_shape must be provided when grid is None")
        grid = grid_factory(grid_shape, compute_backend=compute_backend or DefaultConfig.default_backend)
    q, store = velocity_set.q, precision_policy.store_precision
    layout = ((q, store), (q, store), (q, Precision.BOOL), (1, Precision.UINT8))  # f_0, f_1, missing_mask, bc_mask
    return (grid, *(grid.create_field(cardinality=card, dtype=dtype) for card, dtype in layout))
