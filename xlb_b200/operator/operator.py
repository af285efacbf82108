"""Operator base class and backend dispatch.

Mirrors /root/reference/xlb/operator/operator.py:10-134 — this is the drop-in boundary: users
construct operator objects and call them; ``__call__`` looks up the implementations registered for
``(class name, compute_backend)`` with ``Operator.register_backend``, binds the arguments against each signature and
runs the first that fits, wrapping any failure into a generic ``Exception`` with the traceback text (L54-74).

Both enum members route to the native CUDA library.  The member only selects the reference's CALL CONVENTION:
``ComputeBackend.JAX`` = functional (outputs are allocated and returned), ``ComputeBackend.WARP`` = output buffers are
passed in, filled and returned.
"""

import inspect
import traceback

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.default_config import DefaultConfig


class Operator:
    _backends = {}

    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None):
        self.velocity_set = velocity_set or DefaultConfig.velocity_set
        self.precision_policy = precision_policy or DefaultConfig.default_precision_policy
        self.compute_backend = compute_backend or DefaultConfig.default_backend
        if self.compute_backend not in ComputeBackend:
            raise ValueError(f"Compute_backend {compute_backend} is not supported")
        if self.velocity_set is None or self.precision_policy is None:
            raise ValueError("velocity_set / precision_policy not given and xlb.init(...) was not called")

    @classmethod
    def register_backend(cls, backend_name):
        """Decorator: register an implementation for a compute backend (reference: operator.py:39-52)."""

        def decorator(func):
            owner = func.__qualname__.split(".")[0]
            cls._backends[(owner, backend_name, str(inspect.signature(func)))] = func
            return func

        return decorator

    def _candidates(self):
        for klass in type(self).__mro__:
            found = [(k, m) for k, m in self._backends.items() if k[0] == klass.__name__ and k[1] == self.compute_backend]
            if found:
                return found
        return []

    def __call__(self, *args, callback=None, **kwargs):
        key, error, traceback_str = None, None, ""
        for key, backend_method in self._candidates():
            try:
                bound = inspect.signature(backend_method).bind(self, *args, **kwargs)
                bound.apply_defaults()
                result = backend_method(self, *args, **kwargs)
                if callback and callable(callback):
                    callback(result if result is not None else (args, kwargs))
                return result
            except Exception as e:  # try the next signature (reference: operator.py:69-72)
                error = e
                traceback_str = traceback.format_exc()
                continue
        raise Exception(f"Error captured for backend with key {key} for operator {self.__class__.__name__}: {error}\n {traceback_str}")

    @property
    def supported_compute_backend(self):
        return list(self._backends.keys())

    def __repr__(self):
        return f"{self.__class__.__name__}()"

    @property
    def compute_dtype(self):
        return self.precision_policy.compute_precision.torch_dtype

    @property
    def store_dtype(self):
        return self.precision_policy.store_precision.torch_dtype

    # -- shared helpers for the native calls ---------------------------------------------------------------------
    @property
    def _lattice(self) -> int:
        return self.velocity_set.lattice_code

    @property
    def _compute_code(self) -> int:
        return self.precision_policy.compute_precision.code
