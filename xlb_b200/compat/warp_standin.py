"""Minimal `warp` stand-in (host-side helpers used by reference example scripts; see xlb_b200/compat/__init__.py).

`@wp.func` profiles are ordinary Python callables here; xlb_b200 evaluates them on the host, vectorised over all
boundary cells (index[k] is then an integer ARRAY), which is why the scalar helpers below are numpy-aware."""

import numpy as np
import torch

__standin__ = True

float16, float32, float64 = np.float16, np.float32, np.float64
int32, int64, uint8, uint32 = np.int32, np.int64, np.uint8, np.uint32
bool = np.bool_


def init(*args, **kwargs):
    return None


def synchronize(*args, **kwargs):
    if torch.cuda.is_available():
        torch.cuda.synchronize()


synchronize_device = synchronize_stream = synchronize


def func(f=None, **kwargs):
    return f if f is not None else (lambda g: g)


def kernel(f=None, **kwargs):
    def _no_launch(*a, **k):
        raise NotImplementedError("user-defined Warp kernels cannot run on the xlb_b200 backend (no Warp compiler)")

    return _no_launch


def launch(*args, **kwargs):
    raise NotImplementedError("wp.launch is not available: xlb_b200 runs its own CUDA kernels")


def constant(x):
    return x


def static(x):
    return x


def vec(*values, length=None, dtype=None):
    """wp.vec(x, length=1) inside profiles -> a plain list the BC code indexes with [0]."""
    if length is not None and len(values) == 1:
        return [values[0]] * int(length)
    return list(values)


def vec3i(*v):
    return tuple(int(x) for x in v) if v else (0, 0, 0)


vec2i = vec3i
vec3 = vec3f = vec2 = vec2f = lambda *v: tuple(float(x) for x in v)


def max(a, b):  # noqa: A001 (mirrors wp.max)
    return np.maximum(a, b)


def min(a, b):  # noqa: A001
    return np.minimum(a, b)


abs = np.abs  # noqa: A001
sqrt, sin, cos, exp, log, floor, ceil, pow = np.sqrt, np.sin, np.cos, np.exp, np.log, np.floor, np.ceil, np.power


def to_jax(a):
    """wp.to_jax(field): the same memory seen as a JAX-convention array."""
    from xlb_b200.field import Field, WarpField

    return a.as_subclass(Field) if isinstance(a, WarpField) else a


def from_jax(a, dtype=None):
    return a


to_torch = from_torch = to_jax


def copy(dst, src):
    dst.copy_(src)
    return dst


def clone(a):
    return a.clone()


def _torch_dtype(dtype):
    return {np.float16: torch.float16, np.float32: torch.float32, np.float64: torch.float64, np.uint8: torch.uint8, np.int32: torch.int32, np.bool_: torch.bool}.get(dtype, dtype)


def zeros(shape, dtype=float32, device=None):
    from xlb_b200.field import Field
    from xlb_b200.grid.grid import default_device

    return Field.wrap(torch.zeros(shape, dtype=_torch_dtype(dtype), device=default_device()))


def full(shape, value, dtype=float32, device=None):
    from xlb_b200.field import Field
    from xlb_b200.grid.grid import default_device

    return Field.wrap(torch.full(shape, value, dtype=_torch_dtype(dtype), device=default_device()))


def array(data, dtype=None, device=None):
    from xlb_b200.field import as_field
    from xlb_b200.grid.grid import default_device

    return as_field(np.asarray(data), dtype=_torch_dtype(dtype) if dtype is not None else None, device=default_device())


class ScopedTimer:
    def __init__(self, name="", active=True, **kwargs):
        self.name, self.active = name, active

    def __enter__(self):
        import time

        synchronize()
        self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        import time

        synchronize()
        self.elapsed = (time.perf_counter() - self._t0) * 1e3
        if self.active:
            print(f"{self.name} took {self.elapsed:.2f} ms")


class _Config:
    mode = "release"
    verbose = False
    quiet = True


config = _Config()
