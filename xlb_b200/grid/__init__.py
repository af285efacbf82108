"""Grid namespace: one slab-aware Grid class; WarpGrid / JaxGrid only select the field-shape convention."""

from xlb_b200._exports import export

export(globals(), __name__, {"grid": ["grid_factory", "Grid", "WarpGrid", "JaxGrid"]})
