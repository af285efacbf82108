"""Minimal `jax` / `jax.numpy` stand-in for the host-side post-processing lines of reference example scripts
(see xlb_b200/compat/__init__.py).  Arrays are torch tensors (device fields) or numpy arrays; each function dispatches
on what it is given.  Nothing here is used by the LBM step."""

import types

import numpy as _np
import torch as _torch

__standin__ = True


def _is_t(*xs):
    return any(isinstance(x, _torch.Tensor) for x in xs)


def _unary(name, tname=None):
    nf, tf = getattr(_np, name), getattr(_torch, tname or name)
    return lambda x, *a, **k: tf(x, *a, **k) if _is_t(x) else nf(x, *a, **k)


def _binary(name, tname=None):
    nf, tf = getattr(_np, name), getattr(_torch, tname or name)

    def f(a, b):
        if _is_t(a, b):
            a = a if isinstance(a, _torch.Tensor) else _torch.as_tensor(a, device=b.device, dtype=b.dtype)
            b = b if isinstance(b, _torch.Tensor) else _torch.as_tensor(b, device=a.device, dtype=a.dtype)
            return tf(a, b)
        return nf(a, b)

    return f


numpy = types.ModuleType("jax.numpy")
class _JaxArrayMeta(type):
    def __instancecheck__(cls, x):  # device fields are "jax arrays" unless they were created on the WARP convention (a wp.array is not)
        from xlb_b200.field import WarpField

        return isinstance(x, _torch.Tensor) and not isinstance(x, WarpField)


class _JaxArray(metaclass=_JaxArrayMeta):
    pass


numpy.ndarray = _JaxArray
for _n in ("sqrt", "abs", "sin", "cos", "exp", "log", "square", "isnan", "zeros_like", "ones_like"):
    setattr(numpy, _n, _unary(_n))
for _n in ("maximum", "minimum"):
    setattr(numpy, _n, _binary(_n))
numpy.sum = lambda x, axis=None, **k: (x.sum(dim=axis, **k) if axis is not None else x.sum()) if _is_t(x) else _np.sum(x, axis=axis, **k)
numpy.mean = lambda x, axis=None: (x.mean(dim=axis) if axis is not None else x.mean()) if _is_t(x) else _np.mean(x, axis=axis)
numpy.max = lambda x, axis=None: (x.amax(dim=axis) if axis is not None else x.max()) if _is_t(x) else _np.max(x, axis=axis)
numpy.min = lambda x, axis=None: (x.amin(dim=axis) if axis is not None else x.min()) if _is_t(x) else _np.min(x, axis=axis)
numpy.where = lambda c, a=None, b=None: _torch.where(c, a, b) if _is_t(c, a, b) else (_np.where(c) if a is None else _np.where(c, a, b))
numpy.stack = lambda xs, axis=0: _torch.stack(list(xs), dim=axis) if _is_t(*xs) else _np.stack(xs, axis=axis)
numpy.concatenate = lambda xs, axis=0: _torch.cat(list(xs), dim=axis) if _is_t(*xs) else _np.concatenate(xs, axis=axis)
for _n in ("arange", "meshgrid", "linspace", "zeros", "ones", "full", "array", "asarray", "float16", "float32", "float64", "int32", "uint8", "bool_", "pi", "newaxis"):
    setattr(numpy, _n, getattr(_np, _n))


def jit(fun=None, **kwargs):
    return fun if fun is not None else (lambda f: f)


def device_count():
    return _torch.cuda.device_count() if _torch.cuda.is_available() else 1


def devices(*a):
    return [f"cuda:{i}" for i in range(device_count())]


def default_backend():
    return "gpu" if _torch.cuda.is_available() else "cpu"


class _Config:
    @staticmethod
    def update(key, value):
        return None


config = _Config()
