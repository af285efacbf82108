"""TEST INFRASTRUCTURE — stand-in `jax` / `warp` modules so the REFERENCE's own Python can run here.

The reference (Autodesk/XLB, /root/reference) is 100 % Python but imports `jax` and `warp` at module import time;
neither is installed or installable in the build container (no network, no wheel).  All LBM arithmetic of the
reference's JAX backend is written out in the reference's own files as array expressions (`jnp.roll`, `jnp.where`,
`jnp.tensordot`, `x.at[i].set(v)` ...).  This package provides *only those array primitives*, backed by numpy and
following JAX's dtype rules (float32 default, python scalars are weakly typed, int (+) float32 -> float32), plus two
flavours of `warp`: inert placeholders (default; enough to import the package and run the JAX backend), or — with
`install(interpret_warp=True)` — an INTERPRETIVE stand-in that executes the reference's Warp kernels and functionals
as the plain Python they are, cell by cell (see `_make_warp_interp`), so that `ComputeBackend.WARP` operators run too.

`install()` puts the stand-ins into `sys.modules`; afterwards `import xlb` (with /root/reference on sys.path) works and
the reference's operators execute its own code line by line.  Used exclusively by tests/golden/make_golden.py and
make_golden_warp.py to generate golden vectors (committed as .npz) and by tests that are skipped when /root/reference is
absent.  Nothing in the product path imports this.
"""

from __future__ import annotations

import sys
import types

import numpy as np

_X64 = {"enabled": False}


def _default_float():
    return np.float64 if _X64["enabled"] else np.float32


# ------------------------------------------------------------------------------------------------
# array type with `.at[idx].set(value)` and JAX-like type promotion
# ------------------------------------------------------------------------------------------------


class _At:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self._arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self._arr, self._idx = arr, idx

    def set(self, value):
        out = np.array(self._arr, copy=True).view(JArray)
        out[self._idx] = np.asarray(value)
        return out


def _float_rank(dt):
    return {np.dtype(np.float16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}.get(np.dtype(dt), 0)


def _promote_inputs(inputs):
    """JAX rule used by the reference's expressions: integer/bool arrays combined with a floating array take the
    floating dtype (numpy would go to float64); numpy float64 *scalars* are treated as weak python floats."""
    target, rank = None, 0
    for x in inputs:
        if isinstance(x, np.ndarray) and x.ndim > 0 and _float_rank(x.dtype) > rank:
            target, rank = x.dtype, _float_rank(x.dtype)
    if target is None:
        return inputs
    out = []
    for x in inputs:
        if isinstance(x, np.ndarray) and x.ndim > 0 and x.dtype.kind in "iu":
            x = x.astype(target)
        elif isinstance(x, (np.floating,)) or (isinstance(x, np.ndarray) and x.ndim == 0 and x.dtype.kind == "f"):
            x = float(x) if _float_rank(getattr(x, "dtype", np.float64)) > rank else x
        out.append(x)
    return tuple(out)


def _finish(x):
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64 and not _X64["enabled"]:
            x = x.astype(np.float32)
        return x.view(JArray)
    if isinstance(x, tuple):
        return tuple(_finish(v) for v in x)
    if isinstance(x, list):
        return [_finish(v) for v in x]
    return x


class JArray(np.ndarray):
    @property
    def at(self):
        return _At(self)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        raw = tuple(np.asarray(x) if isinstance(x, JArray) else x for x in inputs)
        raw = _promote_inputs(raw)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, JArray) else o for o in out)
        res = getattr(ufunc, method)(*raw, **kwargs)
        return _finish(res)

    def astype(self, dtype, *a, **k):
        return np.asarray(self).astype(dtype, *a, **k).view(JArray)

    def block_until_ready(self):
        return self


# ------------------------------------------------------------------------------------------------
# jax.numpy
# ------------------------------------------------------------------------------------------------


def _unwrap(x):
    if isinstance(x, JArray):
        return np.asarray(x)
    if isinstance(x, (tuple, list)):
        return type(x)(_unwrap(v) for v in x)
    return x


def _wrap_fn(name):
    fn = getattr(np, name)

    def wrapped(*args, **kwargs):
        args = _promote_inputs(tuple(_unwrap(a) for a in args))
        kwargs = {k: _unwrap(v) for k, v in kwargs.items()}
        if "axis" in kwargs and hasattr(kwargs["axis"], "__next__"):  # jax accepts any iterable of axes
            kwargs["axis"] = tuple(kwargs["axis"])
        return _finish(fn(*args, **kwargs))

    wrapped.__name__ = name
    return wrapped


def _jnp_array(obj, dtype=None, copy=True):
    a = np.array(_unwrap(obj), dtype=dtype, copy=True)
    if dtype is None:
        if a.dtype == np.float64 and not _X64["enabled"]:
            a = a.astype(np.float32)
        elif a.dtype == np.int64 and not _X64["enabled"]:
            a = a.astype(np.int32)
    return a.view(JArray)


def _creation(name):
    fn = getattr(np, name)

    def wrapped(shape, *args, dtype=None, **kwargs):
        if dtype is None and name in ("zeros", "ones", "empty"):
            dtype = _default_float()
        return fn(shape, *args, dtype=dtype, **kwargs).view(JArray)

    return wrapped


def _jnp_full(shape, fill_value, dtype=None):
    if dtype is None:
        dtype = _default_float() if isinstance(fill_value, float) else None
    return np.full(shape, fill_value, dtype=dtype).view(JArray)


def _jnp_sqrt(x):
    if isinstance(x, (int, float)):
        return _default_float()(np.sqrt(_default_float()(x)))  # jax: computed and typed in the default float type
    return _finish(np.sqrt(_unwrap(x)))


def _jnp_tensordot(a, b, axes=2):
    a, b = _promote_inputs((_unwrap(a), _unwrap(b)))
    return _finish(np.tensordot(a, b, axes=axes))


def _jnp_where(cond, x=None, y=None):
    if x is None:
        return tuple(v.view(JArray) for v in np.where(_unwrap(cond)))
    x, y = _unwrap(x), _unwrap(y)
    xs = [v for v in (x, y) if isinstance(v, np.ndarray) and v.ndim > 0]
    res = np.where(_unwrap(cond), x, y)
    if xs and all(v.dtype == xs[0].dtype for v in xs):
        res = res.astype(xs[0].dtype, copy=False)
    return _finish(res)


def _jnp_pad(array, pad_width, mode="constant", **kwargs):
    return _finish(np.pad(_unwrap(array), pad_width, mode=mode, **kwargs))


def _make_jnp():
    m = types.ModuleType("jax.numpy")
    m.ndarray = np.ndarray  # `isinstance(x, jnp.ndarray)` is true for our arrays
    m.array = _jnp_array
    m.asarray = lambda obj, dtype=None: _jnp_array(obj, dtype=dtype)
    for name in ("zeros", "ones", "empty"):
        setattr(m, name, _creation(name))
    m.full = _jnp_full
    m.sqrt = _jnp_sqrt
    m.tensordot = _jnp_tensordot
    m.where = _jnp_where
    m.pad = _jnp_pad
    for name in (
        "zeros_like", "ones_like", "sum", "square", "roll", "logical_and", "logical_or", "logical_not", "broadcast_to",
        "stack", "arange", "meshgrid", "maximum", "minimum", "abs", "sin", "cos", "rint", "dot", "concatenate",
        "reshape", "transpose", "mean", "max", "min", "linspace", "expand_dims", "squeeze", "any", "all", "isnan",
    ):  # fmt: skip
        setattr(m, name, _wrap_fn(name))
    for name in ("float16", "float32", "float64", "int32", "int64", "uint8", "bool_", "pi", "newaxis", "inf", "nan"):
        setattr(m, name, getattr(np, name))
    return m


# ------------------------------------------------------------------------------------------------
# jax, jax.lax, sharding placeholders
# ------------------------------------------------------------------------------------------------


def _jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def _vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        n = len(args[0])
        outs = [fun(*[a[i] for a in args]) for i in range(n)]
        return _finish(np.stack([np.asarray(o) for o in outs], axis=out_axes))

    return mapped


def _broadcast_in_dim(operand, shape, broadcast_dimensions):
    operand = np.asarray(operand)
    view = [1] * len(shape)
    for src, dst in enumerate(broadcast_dimensions):
        view[dst] = operand.shape[src]
    return np.broadcast_to(operand.reshape(view), shape).view(JArray)


class _Device:
    platform = "cpu"
    id = 0

    def __repr__(self):
        return "CpuDevice(shim)"


_DEVICE = _Device()


class _Mesh:
    def __init__(self, devices, axis_names=None):
        self.devices, self.axis_names = devices, axis_names


class _PartitionSpec(tuple):
    def __new__(cls, *args):
        return super().__new__(cls, args)


class _NamedSharding:
    def __init__(self, mesh, spec):
        self.mesh, self.spec = mesh, spec

    def addressable_devices_indices_map(self, shape):
        return {_DEVICE: tuple(slice(None) for _ in shape)}


class _Permissive(types.ModuleType):
    """Module whose unknown attributes are inert callables/types (for never-executed Warp/plotting code)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _Inert(f"{self.__name__}.{name}")
        setattr(self, name, obj)
        return obj


class _Inert:
    def __init__(self, name="inert"):
        self._name = name

    def __call__(self, *a, **k):
        # decorator use (@wp.func / @wp.kernel): hand the function back untouched
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Inert):
            return a[0]
        return _Inert(self._name + "()")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert(self._name + "." + name)

    def __getitem__(self, k):
        return self

    def __mro_entries__(self, bases):
        return (object,)

    def __repr__(self):
        return f"<inert {self._name}>"


def _make_warp():
    wp = _Permissive("warp")
    wp.constant = lambda x: x
    wp.init = lambda *a, **k: None
    wp.synchronize = lambda *a, **k: None

    def _vec(*a, **k):
        return lambda *vals: np.asarray(vals[0]) if vals else None

    wp.vec = _vec
    wp.mat = _vec
    for n, t in (("float16", np.float16), ("float32", np.float32), ("float64", np.float64), ("uint8", np.uint8), ("int32", np.int32), ("bool", np.bool_)):
        setattr(wp, n, t)
    wp.sqrt = lambda x: float(np.sqrt(x))
    utils = _Permissive("warp.utils")
    wp.utils = utils
    return wp, utils


# ---- interpretive `warp`: executes the reference's Warp kernels / functionals as the plain Python they are ---------------
# Warp kernels are syntactically Python.  With small value-type vector classes (numpy-backed), `wp.func` = call with
# by-value vector arguments, `wp.kernel` + `wp.launch` = a loop over the launch grid with `wp.tid()` returning the current
# index, the reference's WARP backend (fused step kernel nse_stepper.py:344-381, every BC functional, the Warp masker, the
# aux-data kernels) runs here cell by cell — slowly, but it IS the reference's code, and it pins the Warp-only semantics
# (255 skip, scalar prescribed value kept in f_1[0], aux recovery, per-index interior flag of the masker) that the JAX
# path cannot.  Typing follows Warp's: numpy scalars keep their width (NEP 50 makes Python literals weak, as in Warp),
# Python floats handed to a launch become float32 (Warp's inference for `Any`-typed scalar arguments).
class _VecBase(np.ndarray):
    pass


class WpArray(np.ndarray):
    """numpy array with the two wp.array members the reference touches."""

    def numpy(self):
        return np.array(self, copy=True).view(np.ndarray)

    @property
    def ptr(self):
        return self.ctypes.data


_VEC_TYPES = {}


def _vec_type(n, dtype):
    key = (int(n), np.dtype(dtype))
    if key not in _VEC_TYPES:
        length, dt = key

        class Vec(_VecBase):
            def __new__(cls, *vals):
                a = np.zeros(length, dtype=dt).view(cls)
                if len(vals) == 1:
                    a[:] = np.asarray(vals[0]).reshape(-1) if np.ndim(vals[0]) else vals[0]
                elif vals:
                    a[:] = vals
                return a

        Vec._length, Vec._scalar = length, dt
        Vec.__name__ = f"vec{length}_{dt.name}"
        _VEC_TYPES[key] = Vec
    return _VEC_TYPES[key]


def _mat_type(shape, dtype):
    dt = np.dtype(dtype)

    def make(values=None):
        a = np.zeros(shape, dtype=dt)
        if values is not None:
            a[...] = np.asarray(values).reshape(shape)
        return a

    return make


_TID = {"idx": None}


class _ReadsFirst:
    """One launch's view of an array: reads see the PRE-LAUNCH contents, writes go to the live array.

    The reference's fused step kernel has a benign data race: while every thread pulls its neighbours' populations from
    f_0, the aux-recovery step of Zou-He / Regularized cells writes f_1's old values back into f_0[opp[l], cell] for the
    cell's missing directions (nse_stepper.py:318-342) — slots that the periodic-wrap neighbour on the opposite domain
    face (a wall-edge cell) pulls in the same launch.  The raced values only ever reach cells behind a boundary (they
    are bounced straight back out of the domain), but they are visible in the arrays.  A plain sequential loop would
    resolve the race by launch-index order; this view resolves it as reads-before-writes, the one ordering whose result
    does not depend on thread scheduling (no thread of a launch reads a slot it wrote itself, except through atomic_add,
    which goes to the live array)."""

    __slots__ = ("live", "snap", "shape", "dtype", "ndim")

    def __init__(self, live):
        self.live, self.snap = live, np.array(live, copy=True).view(np.ndarray)
        self.shape, self.dtype, self.ndim = live.shape, live.dtype, live.ndim

    def __getitem__(self, k):
        return self.snap[k]

    def __setitem__(self, k, v):
        self.live[k] = v


class _Kernel:
    def __init__(self, fn):
        self.fn = fn
        self.__name__ = getattr(fn, "__name__", "kernel")


def _make_warp_interp():
    import functools

    wp = _Permissive("warp")
    wp.__interp__ = True
    wp.constant = lambda x: x
    wp.static = lambda x: x
    wp.init = lambda *a, **k: None
    wp.synchronize = lambda *a, **k: None
    for n, t in (("float16", np.float16), ("float32", np.float32), ("float64", np.float64), ("uint8", np.uint8), ("int32", np.int32),
                 ("int64", np.int64), ("bool", np.bool_)):  # fmt: skip
        setattr(wp, n, t)

    def vec(*args, length=None, dtype=None):
        if args and length is not None:  # value form: wp.vec(x, length=n)   (bc_zouhe.py:113)
            val = args[0]
            return _vec_type(length, dtype or np.asarray(val).dtype)(val)
        return _vec_type(args[0] if args else length, dtype)

    wp.vec = vec
    wp.mat = lambda shape, dtype=None: _mat_type(tuple(shape), dtype)
    wp.vec2i, wp.vec3i = _vec_type(2, np.int32), _vec_type(3, np.int32)
    wp.vec2, wp.vec3 = _vec_type(2, np.float32), _vec_type(3, np.float32)
    wp.vec2f, wp.vec3f, wp.vec2d, wp.vec3d = wp.vec2, wp.vec3, _vec_type(2, np.float64), _vec_type(3, np.float64)

    def func(fn):
        @functools.wraps(fn)
        def by_value(*a, **k):  # Warp vectors are value types: the callee works on its own copy
            return fn(*(x.copy() if isinstance(x, _VecBase) else x for x in a), **k)

        return by_value

    wp.func = func
    wp.kernel = lambda fn: _Kernel(fn)

    def tid():
        return _TID["idx"]

    wp.tid = tid

    def launch(kernel=None, dim=None, inputs=(), outputs=(), **kw):
        dims = (int(dim),) if np.ndim(dim) == 0 else tuple(int(d) for d in dim)
        args = [np.float32(a) if type(a) is float else (_ReadsFirst(a) if isinstance(a, WpArray) else a) for a in list(inputs) + list(outputs)]
        for idx in np.ndindex(*dims):
            _TID["idx"] = idx[0] if len(idx) == 1 else idx
            kernel.fn(*args)
        _TID["idx"] = None

    wp.launch = launch

    def _as_array(a, dtype=None):
        if isinstance(dtype, type) and issubclass(dtype, _VecBase):  # array of vectors: trailing axis = components
            return np.array(a, dtype=dtype._scalar, copy=True).reshape(-1, dtype._length).view(WpArray)
        return np.array(a, dtype=int if dtype is int else dtype, copy=True).view(WpArray)

    wp.array = lambda data=None, dtype=None, **k: None if data is None else _as_array(data, dtype)
    wp.from_numpy = lambda data, dtype=None, **k: _as_array(data, dtype)
    def zeros(shape, dtype=np.float32, **k):
        shape = (int(shape),) if np.ndim(shape) == 0 else tuple(shape)
        if isinstance(dtype, type) and issubclass(dtype, _VecBase):  # array of vectors (momentum_transfer.py:166)
            return np.zeros(shape + (dtype._length,), dtype=dtype._scalar).view(WpArray)
        return np.zeros(shape, dtype=dtype).view(WpArray)

    wp.zeros = zeros
    wp.empty = zeros
    wp.ones = lambda shape, dtype=np.float32, **k: np.ones(shape, dtype=dtype).view(WpArray)
    wp.full = lambda shape, value, dtype=np.float32, **k: np.full(shape, value, dtype=dtype).view(WpArray)
    wp.clone = lambda a, **k: np.array(a, copy=True).view(WpArray)
    wp.to_jax = lambda a: np.asarray(a).view(np.ndarray)
    wp.from_jax = lambda a, dtype=None: _as_array(a, dtype)

    def copy(dest, src, **k):
        dest[...] = src

    wp.copy = copy
    for n in ("array1d", "array2d", "array3d", "array4d"):
        setattr(wp, n, lambda *a, **k: None)  # annotations only

    def atomic_add(arr, *idx_and_value):
        *idx, value = idx_and_value
        arr = arr.live if isinstance(arr, _ReadsFirst) else arr
        old = arr[tuple(idx)]
        arr[tuple(idx)] = old + value
        return old

    wp.atomic_add = atomic_add
    wp.abs, wp.sqrt, wp.min, wp.max = np.abs, np.sqrt, min, max
    wp.dot = lambda a, b: (a * b).sum(dtype=a.dtype)
    wp.length = lambda a: np.sqrt((a * a).sum(dtype=a.dtype))
    wp.cw_div = lambda a, b: a / b
    wp.cw_mul = lambda a, b: a * b
    # ---- 3x3 matrices and triangle meshes (mesh_boundary_masker.py) --------------------------------------------------
    def mat33(*args):
        """wp.mat33(s) = filled with s; wp.mat33(a, b, c) with three vectors = the vectors as COLUMNS (Warp's convention,
        which is why the reference transposes the result to index vertices / edges by row)."""
        if len(args) == 1 and np.ndim(args[0]) == 0:
            return np.full((3, 3), args[0], dtype=np.float32)
        if len(args) == 3:
            return np.stack([np.asarray(a, dtype=np.float32) for a in args], axis=1)
        return np.asarray(args, dtype=np.float32).reshape(3, 3)

    wp.mat33 = wp.mat33f = mat33
    wp.transpose = lambda m: np.array(m.T, copy=True)
    wp.uint64 = np.uint64
    meshes = {}

    class Mesh:
        """wp.Mesh: points + flat triangle index list.  Queries are brute force over per-triangle bounding boxes."""

        def __init__(self, points, indices, **k):
            self.points = np.asarray(points, dtype=np.float32).reshape(-1, 3)
            self.tris = np.asarray(indices, dtype=np.int64).reshape(-1, 3)
            corners = self.points[self.tris]
            self.lo, self.hi = corners.min(axis=1), corners.max(axis=1)
            self.id = np.uint64(len(meshes) + 1)
            meshes[int(self.id)] = self

    wp.Mesh = Mesh

    def mesh_query_aabb(mesh_id, lower, upper):  # faces whose bounding box overlaps [lower, upper] (inclusive, as Warp's BVH)
        m = meshes[int(mesh_id)]
        hit = np.all(m.lo <= np.asarray(upper), axis=1) & np.all(m.hi >= np.asarray(lower), axis=1)
        return [int(i) for i in np.nonzero(hit)[0]]

    def mesh_eval_position(mesh_id, face, u, v):  # p*u + q*v + r*(1-u-v)
        m = meshes[int(mesh_id)]
        p, q, r = (m.points[i] for i in m.tris[face])
        return (p * np.float32(u) + q * np.float32(v) + r * np.float32(1.0 - u - v)).view(wp.vec3)

    def mesh_eval_face_normal(mesh_id, face):  # normalize(cross(q - p, r - p))
        m = meshes[int(mesh_id)]
        p, q, r = (m.points[i] for i in m.tris[face])
        n = np.cross(q - p, r - p).astype(np.float32)
        length = np.sqrt((n * n).sum(dtype=np.float32))
        return ((n / length) if length > 0 else np.zeros(3, np.float32)).view(wp.vec3)

    wp.mesh_query_aabb, wp.mesh_eval_position, wp.mesh_eval_face_normal = mesh_query_aabb, mesh_eval_position, mesh_eval_face_normal
    utils = _Permissive("warp.utils")
    wp.utils = utils
    return wp, utils


_DUMMY_MODULES = (
    "trimesh", "pyvista", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "matplotlib.cm", "cupy", "stl", "PIL",
    "PIL.Image", "pxr", "kvikio", "kvikio._lib", "kvikio._lib.arr", "mpi4py",
)  # fmt: skip


def install(interpret_warp=False):
    """Register the stand-ins in sys.modules (idempotent).  Refuses to shadow a real jax installation.
    interpret_warp: install the interpretive `warp` (runs the reference's WARP backend per cell) instead of the inert one."""
    if "jax" in sys.modules and not getattr(sys.modules["jax"], "__refshim__", False):
        raise RuntimeError("a real `jax` is already imported; the stand-in must not shadow it")
    if getattr(sys.modules.get("jax"), "__refshim__", False):
        return
    jax = types.ModuleType("jax")
    jax.__refshim__ = True
    jax.__path__ = []
    jnp = _make_jnp()
    lax = types.ModuleType("jax.lax")
    lax.broadcast_in_dim = _broadcast_in_dim

    def _no_ppermute(*a, **k):
        raise NotImplementedError("ppermute needs real multi-device jax; the oracle emulates it (stream_sharded)")

    lax.ppermute = _no_ppermute
    sharding = types.ModuleType("jax.sharding")
    sharding.PartitionSpec, sharding.NamedSharding, sharding.Mesh = _PartitionSpec, _NamedSharding, _Mesh
    experimental = types.ModuleType("jax.experimental")
    experimental.__path__ = []
    mesh_utils = types.ModuleType("jax.experimental.mesh_utils")
    mesh_utils.create_device_mesh = lambda shape, *a, **k: np.full(shape, _DEVICE, dtype=object)
    shard_map = types.ModuleType("jax.experimental.shard_map")
    shard_map.shard_map = lambda f, **k: f
    experimental.mesh_utils, experimental.shard_map = mesh_utils, shard_map
    image = _Permissive("jax.image")
    dlpack = _Permissive("jax.dlpack")

    class _Config:
        @staticmethod
        def update(key, value):
            if key == "jax_enable_x64":
                _X64["enabled"] = bool(value)

    jax.config = _Config()
    jax.jit, jax.vmap = _jit, _vmap
    jax.numpy, jax.lax, jax.sharding, jax.experimental, jax.image, jax.dlpack = jnp, lax, sharding, experimental, image, dlpack
    jax.Array = np.ndarray
    jax.devices = lambda *a: [_DEVICE]
    jax.device_count = lambda: 1
    jax.default_backend = lambda: "cpu"
    jax.device_put = lambda x, d=None: x
    jax.make_array_from_single_device_arrays = lambda shape, sharding, arrays: arrays[0]
    mods = {
        "jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.sharding": sharding, "jax.experimental": experimental,
        "jax.experimental.mesh_utils": mesh_utils, "jax.experimental.shard_map": shard_map, "jax.image": image,
        "jax.dlpack": dlpack,
    }  # fmt: skip
    wp, wp_utils = _make_warp_interp() if interpret_warp else _make_warp()
    mods["warp"], mods["warp.utils"] = wp, wp_utils
    for name in _DUMMY_MODULES:
        if name not in sys.modules:
            mods[name] = _Permissive(name)
            mods[name].__path__ = []
    sys.modules.update(mods)


def set_x64(enabled: bool):
    _X64["enabled"] = bool(enabled)


def import_reference(path="/root/reference", interpret_warp=False):
    """install() + import the reference package `xlb` from `path`; returns the module."""
    install(interpret_warp=interpret_warp)
    if path not in sys.path:
        sys.path.insert(0, path)
    import xlb  # noqa: the reference

    if not xlb.__file__.startswith(path):
        raise RuntimeError(f"`xlb` resolved to {xlb.__file__}, not the reference under {path}")
    return xlb
