"""Operator base class and the x-slab wrapper (namespace of reference xlb/operator)."""

from xlb_b200._exports import export

export(globals(), __name__, {"operator": ["Operator"], "parallel_operator": ["ParallelOperator"]})
