"""Base class of steppers (reference: xlb/operator/stepper/stepper.py:6-34)."""

from xlb_b200.default_config import DefaultConfig
from xlb_b200.operator.operator import Operator


class Stepper(Operator):
    def __init__(self, grid, boundary_conditions):
        self.grid = grid
        self.boundary_conditions = boundary_conditions
        super().__init__(DefaultConfig.velocity_set, DefaultConfig.default_precision_policy, DefaultConfig.default_backend)

    def prepare_fields(self, initializer=None):
        raise NotImplementedError("Subclasses must implement prepare_fields()")
