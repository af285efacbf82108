from xlb_b200.grid.grid import grid_factory as grid_factory
from xlb_b200.grid.grid import Grid, WarpGrid, JaxGrid

__all__ = ["grid_factory", "Grid", "WarpGrid", "JaxGrid"]
