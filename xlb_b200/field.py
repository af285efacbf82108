"""Array container.

torch tensors are the only array container of this framework (device memory,
streams and dtype bookkeeping come from PyTorch; all arithmetic is done by the
CUDA library behind the C ABI).  ``Field`` is a zero-cost ``torch.Tensor``
subclass that adds the two accessors reference scripts use on ``wp.array`` /
``jax.Array`` objects: ``.numpy()`` on a device array
(e.g. reference tests/kernels/collision/test_bgk_collision_warp.py:49-51) and
numpy-style ``.copy()`` (reference nse_stepper.py:89).
"""

import numpy as np
import torch


class Field(torch.Tensor):
    @staticmethod
    def wrap(t: torch.Tensor) -> "Field":
        return t if isinstance(t, Field) else t.as_subclass(Field)

    def numpy(self, *args, **kwargs) -> np.ndarray:  # device arrays are copied to the host
        return self.detach().as_subclass(torch.Tensor).cpu().numpy(*args, **kwargs)

    def copy(self) -> "Field":
        return Field.wrap(self.detach().clone())

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a


class WarpField(Field):
    """A field created on the WARP convention (WarpGrid.create_field).  Same memory, same behaviour; the marker only exists so that a
    reference script's ``isinstance(f, jnp.ndarray)`` is False for it — as it is for a ``wp.array`` — and its ``wp.to_jax(f)`` branch is
    taken (examples/cfd/lid_driven_cavity_2d.py:74-78 drops the trailing singleton axis of 2-D Warp fields there)."""


def as_field(x, dtype=None, device=None) -> Field:
    """Convert numpy / python / tensor input to a contiguous Field."""
    if isinstance(x, torch.Tensor):
        t = x
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if device is not None and t.device != torch.device(device):
            t = t.to(device)
    else:
        t = torch.as_tensor(np.asarray(x), dtype=dtype, device=device)
    return Field.wrap(t.contiguous())
