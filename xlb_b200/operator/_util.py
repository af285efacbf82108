"""Argument normalisation shared by the operator front-ends."""

import numpy as np
import torch

from xlb_b200 import native
from xlb_b200.field import Field
from xlb_b200.grid.grid import default_device


def to_device_field(x, dtype=None, like: torch.Tensor = None) -> Field:
    """Accept torch tensors (must already be on a CUDA device) and host numpy / python input (copied to the device)."""
    if isinstance(x, torch.Tensor):
        t = x
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
    else:
        device = like.device if like is not None else default_device()
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x)), dtype=dtype).to(device)
    native.require_cuda(t.contiguous(), "array")
    return Field.wrap(t.contiguous())


def empty_like_field(ref: torch.Tensor, cardinality: int, dtype) -> Field:
    return Field.wrap(torch.empty((cardinality,) + tuple(ref.shape[1:]), dtype=dtype, device=ref.device))
