"""The library's per-cell algebra, compiled for the HOST from the same source the CUDA kernels inline
(xlb_b200/csrc/lbm_math.cuh with -DXLBN_HOST_MIRROR; harness tests/host_math/mirror.cu), against the numpy oracle.

This is how collision / BC code written while no GPU was available (SmagorinskyLESBGK, ForcedCollision — DESIGN.md §10)
is checked on the CPU: arithmetic and control flow of `collide_cell`, `collide_cell_ext`, `bc_zouhe` are the shipped
ones; only the kernels' memory indexing is not exercised here (the extended step kernels reuse the validated
`step_body` unchanged).  Differences to the oracle come from the explicit fma / reciprocal forms: <= 1e-6 in fp32."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import lbm_numpy as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "mirror.cu")
OUT = os.path.join(HERE, "host_math", "_build", "libmirror.so")
CSRC = os.path.join(os.path.dirname(HERE), "xlb_b200", "csrc")
LATTICE = {"D2Q9": 0, "D3Q19": 1, "D3Q27": 2}
BGK, KBC, SMAG, FORCED = 0, 1, 2, 4
F32, F64 = 1, 2


@pytest.fixture(scope="module")
def mirror():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(CSRC, h) for h in ("lbm_math.cuh", "lattice.cuh", "common.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr[-3000:]
    lib = C.CDLL(OUT)
    lib.mirror_collide.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_double]
    lib.mirror_bc_zouhe.argtypes = [C.c_int, C.c_int, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def random_cells(lat, n, dt, seed=0):
    """Populations [q, n, 1(, 1)] a few percent away from equilibrium."""
    rng = np.random.default_rng(seed)
    sp = (n,) + (1,) * (lat.d - 1)
    rho = (1.0 + 0.02 * rng.standard_normal((1,) + sp)).astype(dt)
    u = (0.05 * rng.standard_normal((lat.d,) + sp)).astype(dt)
    feq = O.equilibrium(rho, u, lat)
    return (feq * (1.0 + 0.03 * rng.standard_normal(feq.shape))).astype(dt)


def cell_major(a):  # [q, n, 1, ...] -> contiguous [n, q]
    return np.ascontiguousarray(a.reshape(a.shape[0], -1).T)


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / np.abs(b).max())


COMBOS = [
    ("D3Q19", BGK), ("D3Q19", BGK | FORCED), ("D3Q19", SMAG), ("D3Q19", SMAG | FORCED),
    ("D3Q27", BGK), ("D3Q27", KBC), ("D3Q27", BGK | FORCED), ("D3Q27", KBC | FORCED), ("D3Q27", SMAG), ("D3Q27", SMAG | FORCED),
    ("D2Q9", BGK), ("D2Q9", KBC), ("D2Q9", BGK | FORCED), ("D2Q9", KBC | FORCED),
]  # fmt: skip


@pytest.mark.parametrize("compute", [F32, F64])
@pytest.mark.parametrize("lattice,coll", COMBOS)
def test_collision_functions_of_the_library_match_the_oracle(mirror, lattice, coll, compute):
    lat = O.Lattice(lattice)
    dt = np.float32 if compute == F32 else np.float64
    f = random_cells(lat, 400, dt)
    omega, smag = 1.83, 0.17
    force = np.array([3e-4, -2e-4, 1e-4][: lat.d] + [0.0] * (3 - lat.d))
    rho, u = O.macroscopic(f, lat)
    feq = O.equilibrium(rho, u, lat)
    base = coll & 3
    want = O.collide_bgk(f, feq, omega) if base == BGK else (O.collide_kbc(f, feq, rho, lat, omega) if base == KBC else O.collide_smagorinsky(f, feq, lat, omega, smag))
    if coll & FORCED:
        want = O.exact_difference_force(want, feq, rho, u, force[: lat.d], lat)
    fin = cell_major(f)
    for fast in ([0, 1] if compute == F32 else [0]):
        out = np.empty_like(fin)
        rc = mirror.mirror_collide(LATTICE[lattice], coll, compute, fast, fin.shape[0], fin.ctypes.data, out.ctypes.data, omega, (C.c_double * 3)(*force), smag)
        assert rc == 0
        # the collision changes f by ~3 %; compare the CHANGE so that an error in the collision term cannot hide behind f itself
        err = rel(out - fin, cell_major(want) - fin)
        assert err <= (5e-6 if compute == F32 else 1e-12), (lattice, coll, fast, err)


def test_unbuilt_combinations_are_rejected(mirror):
    z = np.zeros((1, 27), np.float32)
    assert mirror.mirror_collide(LATTICE["D3Q19"], KBC, F32, 0, 1, z.ctypes.data, z.ctypes.data, 1.0, None, 0.17) == -1  # kbc.py:71-72
    assert mirror.mirror_collide(LATTICE["D2Q9"], SMAG, F32, 0, 1, z.ctypes.data, z.ctypes.data, 1.0, None, 0.17) == -1  # 3-D only


ZOUHE_KINDS = {("zouhe", "velocity"): 5, ("zouhe", "pressure"): 6, ("regularized", "velocity"): 7, ("regularized", "pressure"): 8}


@pytest.mark.parametrize("lattice", ["D2Q9", "D3Q19", "D3Q27"])
@pytest.mark.parametrize("kind,bc_type", list(ZOUHE_KINDS))
def test_zouhe_and_regularized_functionals_match_the_oracle(mirror, lattice, kind, bc_type):
    lat = O.Lattice(lattice)
    dt = np.float32
    n_per_face = 40
    faces = [(a, s) for a in range(lat.d) for s in (-1, 1)]
    n = n_per_face * len(faces)
    f = random_cells(lat, n, dt, seed=5)
    sp = f.shape[1:]
    missing = np.zeros((lat.q,) + sp, bool)
    for i, (a, s) in enumerate(faces):  # a cell on the face whose outward normal is s * e_a misses every l with c_l[a] = -s
        missing[lat.c[a] == -s, i * n_per_face : (i + 1) * n_per_face] = True
    rng = np.random.default_rng(9)
    aux = (0.03 * rng.random(sp) if bc_type == "velocity" else 1.0 + 0.01 * rng.standard_normal(sp)).astype(dt)
    bc = O.BC(kind, 1, np.zeros((lat.d, 1), np.int64), bc_type=bc_type, prescribed=aux)
    bc_mask = np.ones((1,) + sp, np.uint8)
    want = (O.bc_zouhe if kind == "zouhe" else O.bc_regularized)(bc, f, f, bc_mask, missing, lat, flavor="warp")
    bits = sum(missing[l].reshape(-1).astype(np.uint32) << np.uint32(l) for l in range(lat.q)).astype(np.uint32)
    fin, out = cell_major(f), np.empty_like(cell_major(f))
    a = np.ascontiguousarray(aux.reshape(-1))
    assert mirror.mirror_bc_zouhe(LATTICE[lattice], F32, ZOUHE_KINDS[(kind, bc_type)], n, fin.ctypes.data, a.ctypes.data, bits.ctypes.data, out.ctypes.data) == 0
    assert rel(out, cell_major(want)) <= 2e-6
