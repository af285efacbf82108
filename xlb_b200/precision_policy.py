"""Precision and PrecisionPolicy enums.

Semantics follow /root/reference/xlb/precision_policy.py:8-89: a policy is a
(compute, store) pair; populations live in HBM in the *store* dtype and every
kernel converts to the *compute* dtype on load and back (round-to-nearest) on
store.  The dtype properties return torch dtypes (torch tensors are the array
container of this framework); ``wp_dtype`` / ``jax_dtype`` are kept as aliases
because reference tests compare ``field.dtype`` against them
(tests/boundary_conditions/mask/test_bc_indices_masker_warp.py:64-66).
"""

from enum import Enum, auto

import numpy as np
import torch


class Precision(Enum):
    FP64 = auto()
    FP32 = auto()
    FP16 = auto()
    UINT8 = auto()
    BOOL = auto()

    @property
    def torch_dtype(self):
        return _TORCH[self]

    @property
    def np_dtype(self):
        return _NUMPY[self]

    # aliases used by reference scripts/tests
    wp_dtype = torch_dtype
    jax_dtype = torch_dtype

    @property
    def code(self) -> int:
        """dtype code of the C ABI (include/xlb_b200.h, xlbn_dtype)."""
        return _CODE[self]


_TORCH = {
    Precision.FP64: torch.float64,
    Precision.FP32: torch.float32,
    Precision.FP16: torch.float16,
    Precision.UINT8: torch.uint8,
    Precision.BOOL: torch.bool,
}
_NUMPY = {
    Precision.FP64: np.float64,
    Precision.FP32: np.float32,
    Precision.FP16: np.float16,
    Precision.UINT8: np.uint8,
    Precision.BOOL: np.bool_,
}
_CODE = {Precision.FP16: 0, Precision.FP32: 1, Precision.FP64: 2, Precision.UINT8: 3, Precision.BOOL: 4}
_FROM_TORCH = {v: k for k, v in _TORCH.items()}


def precision_of(dtype) -> Precision:
    """Map a torch dtype back to a Precision member."""
    return _FROM_TORCH[dtype]


class PrecisionPolicy(Enum):
    FP64FP64 = auto()
    FP64FP32 = auto()
    FP64FP16 = auto()
    FP32FP32 = auto()
    FP32FP16 = auto()

    @property
    def compute_precision(self) -> Precision:
        return Precision[self.name[:4]]

    @property
    def store_precision(self) -> Precision:
        return Precision[self.name[4:]]

    # reference: precision_policy.py:83-89 (names kept; arrays are torch tensors here)
    def cast_to_compute_jax(self, array):
        return torch.as_tensor(array).to(self.compute_precision.torch_dtype)

    def cast_to_store_jax(self, array):
        return torch.as_tensor(array).to(self.store_precision.torch_dtype)

    cast_to_compute = cast_to_compute_jax
    cast_to_store = cast_to_store_jax
