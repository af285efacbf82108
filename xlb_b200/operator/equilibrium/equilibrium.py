"""Base class of equilibrium operators (reference: xlb/operator/equilibrium/equilibrium.py)."""

from xlb_b200.operator.operator import Operator


class Equilibrium(Operator):
    pass
