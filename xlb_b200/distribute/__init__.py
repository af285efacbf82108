"""x-slab distribution of operators (reference namespace xlb/distribute)."""

from xlb_b200._exports import export

export(globals(), __name__, {"distribute": ["distribute", "distribute_operator"]})
