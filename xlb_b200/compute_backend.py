"""Compute-backend enum (mirrors /root/reference/xlb/compute_backend.py:6-8).

Both members are kept so that reference scripts run unchanged.  In this
framework *both* select the native sm_100a CUDA backend; the member only
decides which of the reference's two call conventions an operator accepts
(JAX = functional, WARP = output buffers passed in) and which of the two
reference masker algorithms is reproduced.
"""

from enum import Enum, auto


class ComputeBackend(Enum):
    JAX = auto()
    WARP = auto()
