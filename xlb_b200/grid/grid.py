"""Dense SoA grid with optional 1-D x-slab decomposition.

Reference behaviour kept (xlb/grid/grid.py:9-86, warp_grid.py:17-32,
jax_grid.py:21-59):

* ``grid_factory(shape)`` -> grid; ``grid.create_field(cardinality, dtype,
  fill_value)`` -> array of shape ``(cardinality, nx, ny, nz)``, C-contiguous,
  i.e. struct-of-arrays over the cardinality with **z unit-stride and x
  slowest**; 2-D grids get a trailing singleton axis on the WARP convention
  (``(card, nx, ny, 1)``) and none on the JAX convention.
* ``bounding_box_indices(remove_edges)`` returns the six (four in 2-D) face
  index lists in GLOBAL coordinates, same names and same ordering.

B200 design: there is ONE grid class.  Fields are torch tensors on
``cuda:LOCAL_RANK``.  When ``torch.distributed`` is initialised with
world_size N > 1 the grid is an x-slab decomposition (the reference's only
parallel strategy, ``P(None, "x", ...)`` in xlb/distribute/distribute.py:60):
rank r owns global planes ``[r*nx/N, (r+1)*nx/N)`` and ``create_field``
allocates only the local slab.  Halo planes are *not* part of the fields; they
live in compact ghost buffers owned by the stepper (xlb_b200/distribute/halo.py).
"""

import os
from typing import Tuple

import numpy as np
import torch

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.default_config import DefaultConfig
from xlb_b200.field import Field, WarpField
from xlb_b200.precision_policy import Precision


def _dist_state():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
    return torch.device("cpu")


class Grid:
    def __init__(self, shape: Tuple[int, ...], compute_backend: ComputeBackend = None, device=None, distributed=None):
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        if self.dim not in (2, 3):
            raise ValueError(f"grid must be 2-D or 3-D, got shape {shape}")
        self.compute_backend = compute_backend or DefaultConfig.default_backend or ComputeBackend.WARP
        self.device = torch.device(device) if device is not None else default_device()
        if self.device.type == "cuda" and _dist_state()[1] > 1:
            # one process per GPU: the native library allocates (BC table, ghost planes) and launches on the CURRENT device, and a
            # reference-style script never calls torch.cuda.set_device itself
            torch.cuda.set_device(self.device)

        rank, world = _dist_state() if distributed is None else distributed
        self.rank, self.nDevices = int(rank), int(world)
        if self.shape[0] % self.nDevices != 0:
            raise ValueError(f"nx = {self.shape[0]} is not divisible by the number of slabs {self.nDevices}")
        nx_local = self.shape[0] // self.nDevices
        self.local_shape = (nx_local,) + self.shape[1:]
        self.start_index = (self.rank * nx_local,) + (0,) * (self.dim - 1)

    # -- fields -------------------------------------------------------------------------------
    def _field_shape(self, cardinality):
        shape = self.local_shape
        if self.dim == 2 and self.compute_backend == ComputeBackend.WARP:
            shape = shape + (1,)  # reference: warp_grid.py:26
        return (cardinality,) + shape

    def create_field(self, cardinality: int, dtype: Precision = None, fill_value=None) -> Field:
        if dtype is None:
            dtype = DefaultConfig.default_precision_policy.store_precision
        tdtype = dtype.torch_dtype if isinstance(dtype, Precision) else dtype
        shape = self._field_shape(cardinality)
        if fill_value is None:
            t = torch.zeros(shape, dtype=tdtype, device=self.device)
        else:
            t = torch.full(shape, fill_value, dtype=tdtype, device=self.device)
        return t.as_subclass(WarpField) if self.compute_backend == ComputeBackend.WARP else Field.wrap(t)

    # -- index helpers ------------------------------------------------------------------------
    def bounding_box_indices(self, remove_edges: bool = False):
        """Face index lists, global coordinates (reference: grid.py:34-86)."""
        lo = 1 if remove_edges else 0
        rng = [np.arange(lo, n - lo if remove_edges else n) for n in self.shape]

        def face(axis, value):
            axes = [rng[a] if a != axis else np.array([value]) for a in range(self.dim)]
            mesh = np.meshgrid(*axes, indexing="ij")
            return [m.reshape(-1).tolist() for m in mesh]

        if self.dim == 2:
            nx, ny = self.shape
            return {"bottom": face(1, 0), "top": face(1, ny - 1), "left": face(0, 0), "right": face(0, nx - 1)}
        nx, ny, nz = self.shape
        return {
            "bottom": face(2, 0),
            "top": face(2, nz - 1),
            "left": face(0, 0),
            "right": face(0, nx - 1),
            "front": face(1, 0),
            "back": face(1, ny - 1),
        }


class WarpGrid(Grid):
    def __init__(self, shape, **kw):
        super().__init__(shape, ComputeBackend.WARP, **kw)


class JaxGrid(Grid):
    def __init__(self, shape, **kw):
        super().__init__(shape, ComputeBackend.JAX, **kw)


def grid_factory(shape: Tuple[int, ...], compute_backend: ComputeBackend = None, **kw) -> Grid:
    compute_backend = compute_backend or DefaultConfig.default_backend
    if compute_backend == ComputeBackend.WARP:
        return WarpGrid(shape, **kw)
    if compute_backend == ComputeBackend.JAX:
        return JaxGrid(shape, **kw)
    raise ValueError(f"Compute backend {compute_backend} is not supported")
