"""ctypes binding of the C ABI in include/xlb_b200.h (libxlb_b200.so, hand-written CUDA for sm_100a).

This is the ONLY compute path of the package: there is no CPU or eager-PyTorch fallback.  If the shared library is
missing, or an operator is handed a tensor that does not live on a CUDA device, the call fails loudly.
PyTorch supplies device memory (`tensor.data_ptr()`), the current stream and `torch.distributed`; nothing else.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XLB_B200_LIB") or os.path.join(_HERE, "libxlb_b200.so")  # env override: tuning builds only
CSRC_DIR = os.path.join(_HERE, "csrc")

# enums of include/xlb_b200.h
D2Q9, D3Q19, D3Q27 = 0, 1, 2
F16, F32, F64, U8, BOOL = 0, 1, 2, 3, 4
BGK, KBC, SMAGORINSKY_LES_BGK, COLLISION_FORCED = 0, 1, 2, 4
MASK_WARP, MASK_JAX = 0, 1
(
    BC_NONE,
    BC_EQUILIBRIUM,
    BC_DO_NOTHING,
    BC_HALFWAY_BOUNCE_BACK,
    BC_FULLWAY_BOUNCE_BACK,
    BC_ZOUHE_VELOCITY,
    BC_ZOUHE_PRESSURE,
    BC_REGULARIZED_VELOCITY,
    BC_REGULARIZED_PRESSURE,
    BC_EXTRAPOLATION_OUTFLOW,
) = range(10)

_DTYPE_CODE = {torch.float16: F16, torch.float32: F32, torch.float64: F64, torch.uint8: U8, torch.bool: BOOL}


class XLBNativeError(RuntimeError):
    """Raised when a C-ABI call returns non-zero; carries the library's error string."""

    def __init__(self, code, message):
        super().__init__(f"xlb_b200 native error {code}: {message}")
        self.code = code


class BcDesc(C.Structure):
    _fields_ = [("id", C.c_int32), ("kind", C.c_int32), ("rho", C.c_double), ("u", C.c_double * 3)]


class StepperDesc(C.Structure):
    _fields_ = [
        ("lattice", C.c_int32),
        ("collision", C.c_int32),
        ("compute_dtype", C.c_int32),
        ("store_dtype", C.c_int32),
        ("n_bc", C.c_int32),
        ("cells_per_thread", C.c_int32),
        ("bcs", C.POINTER(BcDesc)),
    ]


class Domain(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("x_begin", C.c_int32), ("x_count", C.c_int32)]


Int3 = C.c_int32 * 3

# name -> argtypes (restype is int unless listed in _RESTYPE); every symbol declared in include/xlb_b200.h
_P, _I, _LL, _D = C.c_void_p, C.c_int, C.c_longlong, C.c_double
SIGNATURES = {
    "xlbn_version": [],
    "xlbn_last_error": [],
    "xlbn_lattice_tables": [_I, _P, _P, _P],
    "xlbn_stepper_create": [C.POINTER(StepperDesc), C.POINTER(_P)],
    "xlbn_stepper_destroy": [_P],
    "xlbn_stepper_set_force": [_P, C.POINTER(C.c_double)],
    "xlbn_stepper_set_smagorinsky": [_P, _D],
    "xlbn_stepper_prepare": [_P, _D, _P],
    "xlbn_step": [_P, _P, _P, _P, _P, C.POINTER(Domain), _D, _I, _P, _P],
    "xlbn_mask_indices": [_I, _I, _P, _LL, _I, _I, Int3, Int3, Int3, _P, _P, _P, _P],
    "xlbn_mask_finalize_jax": [_I, Int3, Int3, Int3, _P, _P, _P, _P],
    "xlbn_mask_mesh": [_I, _P, _LL, _I, _I, Int3, _P, _P, _P, _P],
    "xlbn_pack_missing": [_I, _P, _P, _LL, _P],
    "xlbn_stream": [_I, _P, _P, _I, Int3, _P],
    "xlbn_equilibrium": [_I, _I, _P, _I, _P, _I, _P, _I, Int3, _P],
    "xlbn_macroscopic": [_I, _I, _P, _I, _P, _I, _P, _I, Int3, _P],
    "xlbn_first_moment": [_I, _I, _P, _I, _P, _I, _P, _I, Int3, _P],
    "xlbn_second_moment": [_I, _I, _P, _I, _P, _I, Int3, _P],
    "xlbn_collide": [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _D, Int3, _P],
    "xlbn_collide_ext": [_I, _I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _D, C.POINTER(C.c_double), _D, Int3, _P],
    "xlbn_exact_difference": [_I, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, C.POINTER(C.c_double), Int3, _P],
    "xlbn_bc_apply": [_I, _I, C.POINTER(BcDesc), _P, _P, _I, _P, _P, Int3, _P],
    "xlbn_momentum_transfer": [_I, _I, C.POINTER(BcDesc), _P, _P, _I, _P, _P, Int3, _P, _P],
    "xlbn_halo_create": [_I, _I, _I, _I, C.POINTER(_P)],
    "xlbn_halo_destroy": [_P],
    "xlbn_halo_export": [_P, C.c_char_p],
    "xlbn_halo_connect": [_P, C.c_char_p, C.c_char_p, _I],
    "xlbn_halo_push": [_P, _P, C.POINTER(Domain), _I, _P],
    "xlbn_halo_signal": [_P, _I, _P],
    "xlbn_halo_wait": [_P, _I, _P],
    "xlbn_halo_set_timeout": [_P, _D],
    "xlbn_halo_timed_out": [_P],
    "xlbn_halo_ghost_ptr": [_P, C.POINTER(_P), C.POINTER(_LL)],
}
_RESTYPE = {"xlbn_last_error": C.c_char_p}


def build(verbose: bool = False, jobs: int = None) -> str:
    """Compile the CUDA library in-tree with nvcc for sm_100a (works without a GPU). Returns the .so path."""
    jobs = jobs or os.cpu_count() or 4
    proc = subprocess.run(["make", "-C", CSRC_DIR, f"-j{jobs}"], capture_output=not verbose, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libxlb_b200.so failed:\n" + (proc.stdout or "") + (proc.stderr or ""))
    lib.cache_clear()
    return LIB_PATH


@lru_cache(maxsize=1)
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found. xlb_b200 has no CPU / PyTorch fallback: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C xlb_b200/csrc)."
        )
    handle = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError here = header / library mismatch
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    return handle


def check(code: int):
    if code != 0:
        raise XLBNativeError(code, lib().xlbn_last_error().decode("utf-8", "replace"))


def dtype_code(dtype) -> int:
    try:
        return _DTYPE_CODE[dtype]
    except KeyError:
        raise TypeError(f"dtype {dtype} is not supported by the native library") from None


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} lives on '{t.device}'. xlb_b200 computes only with its CUDA kernels on a B200 (sm_100a); "
            "there is no CPU fallback."
        )
    if not t.is_contiguous():
        raise ValueError(f"{name} must be C-contiguous [card, nx, ny, nz]")
    return t


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_of(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def int3(values) -> Int3:
    v = list(values)
    return Int3(*(int(x) for x in v))


def dims_of(field: torch.Tensor, d: int):
    """(nx, ny, nz) of a [card, nx, ny(, nz | 1)] field; 2-D fields give nz = 1."""
    shape = tuple(field.shape[1:])
    if d == 2:
        if len(shape) == 3 and shape[2] != 1:
            raise ValueError(f"2-D field with trailing extent {shape[2]}")
        return (shape[0], shape[1], 1)
    if len(shape) != 3:
        raise ValueError(f"3-D field expected, got shape {tuple(field.shape)}")
    return shape


def lattice_tables(lattice: int):
    """(c[3][q], w[q], opp[q]) as compiled into the kernels."""
    import numpy as np

    q = {D2Q9: 9, D3Q19: 19, D3Q27: 27}[lattice]
    c = np.zeros((3, q), dtype=np.int32)
    w = np.zeros(q, dtype=np.float64)
    opp = np.zeros(q, dtype=np.int32)
    got = lib().xlbn_lattice_tables(lattice, c.ctypes.data, w.ctypes.data, opp.ctypes.data)
    if got != q:
        check(got)
    return c, w, opp
