"""Process-global defaults set by ``xlb.init`` (reference: xlb/default_config.py:7-26).

Every operator falls back to these when its ctor arguments are ``None``
(reference: xlb/operator/operator.py:19-23).
"""

from dataclasses import dataclass

from xlb_b200.compute_backend import ComputeBackend


@dataclass
class DefaultConfig:
    default_precision_policy = None
    velocity_set = None
    default_backend = None


def init(velocity_set, default_backend, default_precision_policy):
    if default_backend not in (ComputeBackend.JAX, ComputeBackend.WARP):
        raise ValueError(f"Unsupported compute backend: {default_backend}")
    DefaultConfig.velocity_set = velocity_set
    DefaultConfig.default_backend = default_backend
    DefaultConfig.default_precision_policy = default_precision_policy


def default_backend() -> ComputeBackend:
    return DefaultConfig.default_backend
