"""TEST INFRASTRUCTURE — ctypes wrapper of the C/OpenMP oracle (oracle/lbm_ref.c).

Same role and same rules as oracle/lbm_numpy.py (only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use
it).  It is a per-cell restatement of the reference's fused Warp kernel, fast enough for full-size parity runs
(C1: 128^3 x 1000 steps) and used as the multi-threaded CPU baseline."""

import ctypes as C
import os
import time

import numpy as np

from oracle import lbm_numpy as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblbm_ref.so")

KIND = {"equilibrium": 1, "donothing": 2, "halfway": 3, "fullway": 4, "outflow": 9}
_ZOUHE = {("zouhe", "velocity"): 5, ("zouhe", "pressure"): 6, ("regularized", "velocity"): 7, ("regularized", "pressure"): 8}
_STORE = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}


class LbmDesc(C.Structure):
    _fields_ = [
        ("d", C.c_int), ("q", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("collision", C.c_int), ("compute", C.c_int),
        ("store", C.c_int), ("omega", C.c_double), ("c", C.c_int * 81), ("opp", C.c_int * 27), ("w", C.c_double * 27), ("cc", C.c_double * 162),
        ("qi", C.c_double * 162), ("bc_kind", C.c_int * 256), ("bc_rho", C.c_double * 256), ("bc_u", C.c_double * 768),
        ("has_force", C.c_int), ("force", C.c_double * 3), ("smagorinsky", C.c_double),
    ]  # fmt: skip


def available() -> bool:
    return os.path.exists(_LIB)


def _lib():
    lib = C.CDLL(_LIB)
    assert lib.lbm_ref_sizeof_desc() == C.sizeof(LbmDesc)
    lib.lbm_ref_run.argtypes = [C.POINTER(LbmDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return lib


def make_desc(lat: O.Lattice, shape, policy, collision, omega, bcs, force=None, smagorinsky=0.17):
    cdt, sdt = O.policy_dtypes(policy)
    d = LbmDesc()
    d.d, d.q = lat.d, lat.q
    sh = tuple(shape) + (1,) * (3 - len(shape))
    d.nx, d.ny, d.nz = sh
    d.collision = {"BGK": 0, "KBC": 1, "SmagorinskyLESBGK": 2}[collision]
    d.smagorinsky = float(smagorinsky)
    d.has_force = int(force is not None)
    for a in range(lat.d if force is not None else 0):
        d.force[a] = float(force[a])
    d.compute = 1 if cdt == np.float32 else 2
    d.store = _STORE[np.dtype(sdt)]
    d.omega = float(omega)
    for l in range(lat.q):
        for a in range(lat.d):
            d.c[a * 27 + l] = int(lat.c[a, l])
        d.opp[l] = int(lat.opp[l])
        d.w[l] = float(lat.w[l])
        for t in range(lat.cc.shape[1]):
            d.cc[l * 6 + t] = float(lat.cc[l, t])
            d.qi[l * 6 + t] = float(lat.qi[l, t])
    for bc in bcs:
        kind = _ZOUHE[(bc.kind, bc.bc_type)] if bc.kind in ("zouhe", "regularized") else KIND[bc.kind]
        d.bc_kind[bc.id] = kind
        d.bc_rho[bc.id] = float(bc.rho)
        for a in range(lat.d):
            d.bc_u[bc.id * 3 + a] = float(bc.u[a])
    return d


def write_aux(f1, bcs, bc_mask, missing, lat, policy):
    """Warp-style aux init (boundary_condition.py:119-175): prescribed scalar of Zou-He / Regularized cells into
    f1[0, cell], in the store dtype.  JAX-convention vectors are projected on the outward normal."""
    for bc in bcs:
        if bc.kind not in ("zouhe", "regularized"):
            continue
        cells = np.nonzero(bc_mask[0] == bc.id)
        if bc.bc_type == "pressure":
            val = np.broadcast_to(np.asarray(bc.prescribed, dtype=np.float64), bc_mask[0].shape)[cells]
        else:
            normals = O._normals(missing, lat, "warp")
            vec = O._broadcast_prescribed(np.asarray(bc.prescribed, dtype=np.float64), (lat.d,) + bc_mask.shape[1:])
            val = -(vec * normals).sum(axis=0)[cells]
        f1[(0,) + cells] = val.astype(f1.dtype)


def run(f0, bc_mask, missing, bcs, omega, lat, nsteps, policy="FP32FP32", collision="BGK", threads=None, force=None, smagorinsky=0.17):
    """Same contract as oracle.lbm_numpy.run (user loop with buffer swap); returns the final populations."""
    threads = threads or os.cpu_count() or 1
    cdt, sdt = O.policy_dtypes(policy)
    fa = np.ascontiguousarray(f0, dtype=sdt).copy()
    fb = fa.copy()
    write_aux(fb, bcs, bc_mask, missing, lat, policy)
    d = make_desc(lat, f0.shape[1:], policy, collision, omega, bcs, force, smagorinsky)
    bm = np.ascontiguousarray(bc_mask, dtype=np.uint8)
    mm = np.ascontiguousarray(missing).view(np.uint8)
    which = _lib().lbm_ref_run(C.byref(d), fa.ctypes.data, fb.ctypes.data, bm.ctypes.data, mm.ctypes.data, int(nsteps), int(threads))
    return fa if which == 0 else fb


def cavity_case(lattice, n, policy):
    lat = O.Lattice(lattice)
    shape = (n,) * lat.d
    box, box_ne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
    names = ("bottom", "left", "right") + (("front", "back") if lat.d == 3 else ())
    walls = np.unique(np.concatenate([box[k] for k in names], axis=1), axis=-1)
    bcs = [O.BC("equilibrium", 1, box_ne["top"], rho=1.0, u=(0.02, 0.0, 0.0)[: lat.d]), O.BC("fullway", 2, walls)]
    bc_mask, missing = O.build_masks(bcs, shape, lat, flavor="warp")
    return lat, shape, bcs, bc_mask, missing


def time_cavity(lattice, collision, policy, n, steps, threads, periodic=False):
    """MLUPS of `steps` steps of the mlups_3d.py cavity (or a periodic box) at edge n on `threads` host threads."""
    lat, shape, bcs, bc_mask, missing = cavity_case(lattice, n, policy)
    if periodic:
        bcs, bc_mask, missing = [], np.zeros_like(bc_mask), np.zeros_like(missing)
    f = O.initialize_eq(shape, lat, policy)
    run(f, bc_mask, missing, bcs, 1.0, lat, 1, policy, collision, threads)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    run(f, bc_mask, missing, bcs, 1.0, lat, steps, policy, collision, threads)
    return float(np.prod(shape)) * steps / (time.perf_counter() - t0) / 1e6


class Runner:
    """Persistent-buffer form of `run` for timing: the populations stay in two host arrays and `steps(n)` advances them in place
    (no copies inside the timed call).  Same kernel, same swap convention."""

    def __init__(self, f0, bc_mask, missing, bcs, omega, lat, policy="FP32FP32", collision="BGK", threads=None):
        self.threads = threads or os.cpu_count() or 1
        cdt, sdt = O.policy_dtypes(policy)
        self.a = np.ascontiguousarray(f0, dtype=sdt)
        self.b = self.a.copy()
        write_aux(self.b, bcs, bc_mask, missing, lat, policy)
        self.desc = make_desc(lat, f0.shape[1:], policy, collision, omega, bcs)
        self.bm = np.ascontiguousarray(bc_mask, dtype=np.uint8)
        self.mm = np.ascontiguousarray(missing).view(np.uint8)
        self.lib = _lib()

    def steps(self, n):
        """Advance n steps; returns the seconds spent inside the C call."""
        t0 = time.perf_counter()
        which = self.lib.lbm_ref_run(C.byref(self.desc), self.a.ctypes.data, self.b.ctypes.data, self.bm.ctypes.data, self.mm.ctypes.data, int(n), int(self.threads))
        dt = time.perf_counter() - t0
        if which == 1:
            self.a, self.b = self.b, self.a
        return dt

    @property
    def f(self):
        return self.a


def cavity_runner(lattice, collision, policy, n, threads=None, periodic=False):
    """The mlups_3d.py cavity (or a periodic box) at edge n, ready to be stepped; set-up is lean enough for 512^3 (rest-state
    populations are filled per population, no whole-field temporaries)."""
    lat, shape, bcs, bc_mask, missing = cavity_case(lattice, n, policy)
    if periodic:
        bcs, bc_mask, missing = [], np.zeros_like(bc_mask), np.zeros_like(missing)
    cdt, sdt = O.policy_dtypes(policy)
    f = np.empty((lat.q,) + shape, dtype=sdt)
    for l in range(lat.q):  # initialize_eq(rho = 1, u = 0): f_l = w_l  (helper/initializers.py:5-20)
        f[l] = np.asarray(lat.w[l], dtype=cdt).astype(sdt)
    return Runner(f, bc_mask, missing, bcs, 1.0, lat, policy, collision, threads)
