"""The fused step kernel's SOURCE, compiled for the host (tests/host_math/mirror_step.cu, -DXLBN_HOST_MIRROR) and driven by
plain loops, against the reference vectors — both the JAX-backend ones and the WARP-backend ones — on the CPU.

Everything except the launch itself is the shipped code: `fill_step_params` (pointer tables, periodic wrap, ghost
planes), `step_body` (pull addressing, V cells per thread, x-plane classes), `bc_tail` / `bc_cell` (boundary dispatch, aux
recovery), the collision incl. the extended operators, `store_cells`.  This is what stands in for a GPU run of the code
written after the round's GPU budget was spent, and it re-checks the validated kernels on new vectors.
Tolerances are the GPU tests' (tests/common.py:RTOL)."""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from common import LATE_CASES, RTOL, STEP_CASES, WARP_CASES, WARP_CASES_FP16, WARP_CASES_N4, load_golden, oracle_masks, rel_err, tile_case
from oracle import lbm_c
from oracle import lbm_numpy as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_math", "mirror_step.cu")
OUT = os.path.join(HERE, "host_math", "_build", "libmirror_step.so")
CSRC = os.path.join(ROOT, "xlb_b200", "csrc")
LATTICE = {"D2Q9": 0, "D3Q19": 1, "D3Q27": 2}
COLLISION = {"BGK": 0, "KBC": 1, "SmagorinskyLESBGK": 2}
DTYPE = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
# (lattice, collision code) pairs the library instantiates: base | 4 = forced (XLBN_COLLISION_FORCED), | 8 = lean KBC (kLeanKbc)
PARTS = [("D3Q19", 0), ("D3Q19", 4), ("D3Q19", 2), ("D3Q19", 6), ("D3Q27", 0), ("D3Q27", 1), ("D3Q27", 4), ("D3Q27", 5), ("D3Q27", 2), ("D3Q27", 6),
         ("D3Q27", 9), ("D2Q9", 0), ("D2Q9", 1), ("D2Q9", 4), ("D2Q9", 5), ("D2Q9", 9), ("D2Q9X", 0), ("D2Q9X", 1), ("D3Q27", 17), ("D2Q9", 17), ("D3Q27", 13), ("D2Q9", 13)]  # fmt: skip
SYM = lambda lat, coll: f"mirror_step_{'0x' if lat == 'D2Q9X' else LATTICE[lat]}_{coll}"  # noqa: E731


@pytest.fixture(scope="module")
def mirror():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(CSRC, h) for h in ("step_kernel.cuh", "step_tile.cuh", "lbm_math.cuh", "lattice.cuh", "common.cuh", "error.cu")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        build = os.path.dirname(OUT)
        os.makedirs(build, exist_ok=True)
        base = ["nvcc", "-std=c++17", "-O0", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
        jobs = [base + ["-c", os.path.join(CSRC, "error.cu"), "-o", os.path.join(build, "error.o")]]
        objs = [os.path.join(build, "error.o")]
        for i, (lat, coll) in enumerate(PARTS):  # one object per (lattice, collision), compiled in parallel
            obj = os.path.join(build, SYM(lat, coll) + ".o")
            flags = [f"-DMIRROR_LAT={lat}", f"-DMIRROR_TAG=XLBN_{lat.rstrip('X')}", f"-DMIRROR_COLL={coll}", f"-DMIRROR_NAME={SYM(lat, coll)}"]
            jobs.append(base + flags + (["-DMIRROR_DEFINE_ERROR"] if i == 0 else []) + ["-c", SRC, "-o", obj])
            objs.append(obj)
        running, pending = [], list(jobs)
        while pending or running:
            while pending and len(running) < (os.cpu_count() or 4):
                running.append(subprocess.Popen(pending.pop(0), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            p = running.pop(0)
            log = p.communicate()[0]
            assert p.returncode == 0, log[-3000:]
        proc = subprocess.run(["nvcc", "-shared", "-o", OUT] + objs, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr[-3000:]
    lib = C.CDLL(OUT)
    lib.mirror_last_error.restype = C.c_char_p
    P, I, D = C.c_void_p, C.c_int, C.c_double
    for lat, coll in PARTS:
        getattr(lib, SYM(lat, coll)).argtypes = [I, I, I, I, I, P, P, P, P, P, P, P, P, I, I, D, P, D, P, P, P, P]
    return lib


def mirror_run(lib, g, steps=None, v=1, lean_kbc=False, masks=None, slab_axes=False, exact_kbc=False):
    """The user loop (step, swap) through the host-compiled kernel source; masks and aux data from the oracle helpers
    (or `masks` = (lat, bcs, bc_mask, missing) built by the caller)."""
    lat, bcs, bc_mask, missing = masks if masks is not None else oracle_masks(g, "warp")
    cdt, sdt = O.policy_dtypes(g["policy"])
    fa = np.ascontiguousarray(g["f_init"], dtype=sdt).copy()
    fb = fa.copy()
    lbm_c.write_aux(fb, bcs, bc_mask, missing, lat, g["policy"])  # aux_data_init (boundary_condition.py:119-175): prescribed scalar into f_1[0]
    desc = lbm_c.make_desc(lat, g["shape"], g["policy"], g["collision"], g["omega"], bcs)
    kind = np.array(list(desc.bc_kind), dtype=np.int32)  # the C oracle's kind codes are xlbn_bc_kind's
    rho = np.array(list(desc.bc_rho), dtype=np.float64)
    u = np.array(list(desc.bc_u), dtype=np.float64)
    bits = np.ascontiguousarray(sum(missing[l].astype(np.uint32) << np.uint32(l) for l in range(lat.q)).astype(np.uint32))
    bm = np.ascontiguousarray(bc_mask[0])
    shape = g["shape"]
    # 2-D: kernel extents (1, nx, ny); with slab_axes the x-slab axis order (nx, 1, ny) of csrc/lattice.cuh D2Q9X
    dims = (C.c_int32 * 3)(*(((shape[0], 1, shape[1]) if slab_axes else (1,) + tuple(shape)) if lat.d == 2 else tuple(shape)))
    coll = COLLISION[g["collision"]] | (4 if g["force_vector"] is not None else 0) | (8 if lean_kbc else 0) | (16 if exact_kbc else 0)  # csrc/lbm_math.cuh kLeanKbc, kExactKbc
    force = np.zeros(3)
    if g["force_vector"] is not None:
        force[: lat.d] = g["force_vector"]
    for _ in range(g["steps"] if steps is None else steps):
        rc = getattr(lib, SYM(g["lattice"] + ("X" if slab_axes else ""), coll))(LATTICE[g["lattice"]], coll, DTYPE[np.dtype(cdt)], DTYPE[np.dtype(sdt)], v, fa.ctypes.data, fb.ctypes.data, bm.ctypes.data,
                             bits.ctypes.data, kind.ctypes.data, rho.ctypes.data, u.ctypes.data, C.cast(dims, C.c_void_p), 0, dims[0], g["omega"],
                             force.ctypes.data, g["smagorinsky"], None, None, None, None)  # fmt: skip
        assert rc == 0, lib.mirror_last_error().decode()
        fa, fb = fb, fa
    return fa


@pytest.mark.parametrize("v", [102, 202, 203])
@pytest.mark.parametrize("name", ["cavity_d3q19_bgk_fp32", "cavity_d3q19_bgk_fp32fp16", "sphere_d3q19_bgk_fp32fp16", "sphere_d3q27_kbc_fp32", "cavity_d2q9_kbc_fp32",
                                  "warp_tunnel_d3q19_bgk_zouhe_pressure", "periodic_d3q19_bgk_fp32"])  # fmt: skip
def test_pair_paths_match_the_reference_vectors(mirror, name, v):
    """cells_per_thread 102 = packed fp32x2 pair path, 202 = half2-state pair path (the default for FP32FP16 BGK on the GPU):
    two cells per thread, FADD2 / FMUL2 / FFMA2 arithmetic, reciprocal-based divisions; per-half boundary handling with the
    precomputed EquilibriumBC update in the half2 path."""
    g = load_golden(name)
    if v in (202, 203) and not (g["policy"] == "FP32FP16" and g["collision"] == "BGK"):
        pytest.skip("the half2-state path exists for FP32FP16 BGK only")
    if g["shape"][-1] % 2:
        pytest.skip("odd nz: the library falls back to one cell per thread")
    f = mirror_run(mirror, g, v=v)
    assert rel_err(f, g["f_final"]) <= RTOL[g["policy"]]
    assert rel_err(f, mirror_run(mirror, g)) <= (3e-6 if g["policy"] == "FP32FP32" else 1e-3)
    if v == 203:  # the split boundary variant must give the bits of the half2-state path it specialises
        assert np.array_equal(f, mirror_run(mirror, g, v=202))


def mirror_run_slabs(lib, g, n_slabs, steps):
    """The same user loop on `n_slabs` x-slabs, each with its own arrays and ghost planes; the kernel source itself moves the
    outgoing face populations into the neighbours' ghost planes (`out_lo` / `out_hi`), double-buffered by step parity, as
    xlb_b200/distribute/halo.py arranges it between GPUs.  Returns the re-assembled populations."""
    lat, bcs, bc_mask, missing = oracle_masks(g, "warp")
    cdt, sdt = O.policy_dtypes(g["policy"])
    nx, ny, nz = g["shape"]
    assert nx % n_slabs == 0
    h = nx // n_slabs
    fa = np.ascontiguousarray(g["f_init"], dtype=sdt).copy()
    fb = fa.copy()
    lbm_c.write_aux(fb, bcs, bc_mask, missing, lat, g["policy"])
    desc = lbm_c.make_desc(lat, g["shape"], g["policy"], g["collision"], g["omega"], bcs)
    kind, rho, u = np.array(list(desc.bc_kind), dtype=np.int32), np.array(list(desc.bc_rho)), np.array(list(desc.bc_u))
    bits = sum(missing[l].astype(np.uint32) << np.uint32(l) for l in range(lat.q)).astype(np.uint32)
    sl = [slice(r * h, (r + 1) * h) for r in range(n_slabs)]
    A = [np.ascontiguousarray(fa[:, s]) for s in sl]
    B = [np.ascontiguousarray(fb[:, s]) for s in sl]
    BM = [np.ascontiguousarray(bc_mask[0][s]) for s in sl]
    BITS = [np.ascontiguousarray(bits[s]) for s in sl]
    right, left = list(lat.right), list(lat.left)  # c_x = +1 / -1 populations in slot order (lattice.cuh: xdir_slot)
    # ghosts[rank][parity][side]: side 0 = plane "x = -1" (c_x = +1 populations), side 1 = plane "x = h" (c_x = -1 populations)
    ghosts = [[[np.zeros((len(right), ny, nz), sdt) for _ in range(2)] for _ in range(2)] for _ in range(n_slabs)]
    for r in range(n_slabs):  # prime parity 0 from the initial state (halo.py: exchange_ghost_planes)
        ghosts[r][0][0][...] = A[(r - 1) % n_slabs][right, h - 1]
        ghosts[r][0][1][...] = A[(r + 1) % n_slabs][left, 0]
    dims = (C.c_int32 * 3)(h, ny, nz)
    coll = COLLISION[g["collision"]] | (4 if g["force_vector"] is not None else 0)
    force = np.zeros(3)
    if g["force_vector"] is not None:
        force[: lat.d] = g["force_vector"]
    fn = getattr(lib, f"mirror_step_{LATTICE[g['lattice']]}_{coll}")
    for t in range(steps):
        pin, pout = t & 1, (t + 1) & 1
        for r in range(n_slabs):
            hi, lo = (r + 1) % n_slabs, (r - 1) % n_slabs
            rc = fn(LATTICE[g["lattice"]], coll, DTYPE[np.dtype(cdt)], DTYPE[np.dtype(sdt)], 1, A[r].ctypes.data, B[r].ctypes.data, BM[r].ctypes.data,
                    BITS[r].ctypes.data, kind.ctypes.data, rho.ctypes.data, u.ctypes.data, C.cast(dims, C.c_void_p), 0, h, g["omega"], force.ctypes.data,
                    g["smagorinsky"], ghosts[r][pin][0].ctypes.data, ghosts[r][pin][1].ctypes.data, ghosts[lo][pout][1].ctypes.data,
                    ghosts[hi][pout][0].ctypes.data)  # fmt: skip
            assert rc == 0, lib.mirror_last_error().decode()
        A, B = B, A
    return np.concatenate(A, axis=1)


@pytest.mark.parametrize("n_slabs", [2, 4])
@pytest.mark.parametrize("name", ["warp_tunnel_d3q27_kbc_regularized_outflow", "sphere_d3q19_bgk_zouhe_pressure_fp32", "periodic_d3q19_bgk_fp32", "warp_periodic_d3q27_smagorinsky_forced"])
def test_slab_decomposition_through_ghost_planes_is_bit_identical(mirror, name, n_slabs):
    """x-slabs with one ghost plane per side, outgoing face populations stored into the neighbours' ghost planes
    by the step itself — same bits as the undecomposed run (on the GPU: tests/test_native_slab_gpu.py, scripts/mgpu_check.py)."""
    g = load_golden(name)
    if g["shape"][0] % n_slabs:
        pytest.skip("nx not divisible")
    assert np.array_equal(mirror_run_slabs(mirror, g, n_slabs, 6), mirror_run(mirror, g, steps=6))


@pytest.mark.parametrize("name", [n for n in STEP_CASES + WARP_CASES if "kbc" in n])
def test_lean_kbc_variant_matches_the_reference_vectors(mirror, name):
    """cells_per_thread = 301: the register-lean arrangement of the KBC collision (a tuning candidate for round 2) holds the
    same tolerance against the reference and stays within rounding of the default formulation."""
    g = load_golden(name)
    lean = mirror_run(mirror, g, lean_kbc=True)
    assert rel_err(lean, g["f_final"]) <= RTOL[g["policy"]]
    assert rel_err(lean, mirror_run(mirror, g)) <= 3e-6


@pytest.mark.parametrize("name", [n for n in WARP_CASES_N4 if "kbc_forced" in n])
def test_lean_kbc_under_forced_collision_matches_the_reference_vectors(mirror, name):
    """ForcedCollision(KBC): the lean formulation with the ExactDifference term added in its last pass (what cells_per_thread = 0 selects)
    against the reference's WARP-backend vector and against the literal arrangement."""
    g = load_golden(name)
    lean = mirror_run(mirror, g, lean_kbc=True)
    assert rel_err(lean, g["f_final"]) <= RTOL[g["policy"]]
    assert rel_err(lean, mirror_run(mirror, g)) <= 3e-6


def test_wind_tunnel_with_a_voxelised_mesh_body(mirror):
    """examples/cfd/windtunnel_3d.py:66-96 in small: Fullway walls, Regularized inlet, outflow, and a Halfway body given as a
    triangle mesh (masks from the mesh voxeliser: a 255 shell with boundary cells around it) — kernel source vs the C oracle."""
    from test_mesh_masker import load_mesh_case

    shape, lat = (24, 11, 10), O.Lattice("D3Q27")
    box, bne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
    walls = np.unique(np.concatenate([box[k] for k in ("bottom", "top", "front", "back")], axis=1), axis=-1)
    bcs = [O.BC("fullway", 1, walls), O.BC("regularized", 2, bne["left"], bc_type="velocity", prescribed=np.array([0.03, 0.0, 0.0])),
           O.BC("outflow", 3, bne["right"]), O.BC("halfway", 4, np.zeros((3, 0), np.int64))]  # fmt: skip
    bc_mask, missing = O.build_masks(bcs[:3], shape, lat, flavor="warp")
    verts = load_mesh_case("warp_mesh_octahedron_d3q27")["vertices"] + np.array([2.0, 0.0, 0.0])
    bc_mask, missing = O.build_masks_mesh(verts, 4, bc_mask, missing, lat)
    assert (bc_mask == 255).sum() > 50 and (bc_mask == 4).sum() > 100
    g = dict(lattice="D3Q27", policy="FP32FP32", collision="KBC", shape=shape, omega=1.6, steps=20, f_init=O.initialize_eq(shape, lat),
             force_vector=None, smagorinsky=0.17)  # fmt: skip
    f = mirror_run(mirror, g, masks=(lat, bcs, bc_mask, missing))
    ref = lbm_c.run(g["f_init"], bc_mask, missing, bcs, 1.6, lat, 20, "FP32FP32", "KBC")
    assert np.isfinite(f).all() and rel_err(f, ref) <= 1e-5
    solid = bc_mask[0] == 255
    assert np.array_equal(f[:, solid], g["f_init"][:, solid])  # nse_stepper.py:356-358: solid cells are never written


def test_bc_kind_codes_agree_with_the_header():
    """mirror_run hands the C oracle's kind codes to the kernel source: they must be xlbn_bc_kind's."""
    from xlb_b200 import native

    assert (native.BC_EQUILIBRIUM, native.BC_DO_NOTHING, native.BC_HALFWAY_BOUNCE_BACK, native.BC_FULLWAY_BOUNCE_BACK) == (1, 2, 3, 4)
    assert lbm_c.KIND == {"equilibrium": 1, "donothing": 2, "halfway": 3, "fullway": 4, "outflow": native.BC_EXTRAPOLATION_OUTFLOW}
    assert lbm_c._ZOUHE[("zouhe", "velocity")] == native.BC_ZOUHE_VELOCITY and lbm_c._ZOUHE[("regularized", "pressure")] == native.BC_REGULARIZED_PRESSURE


@pytest.mark.parametrize("name", STEP_CASES + LATE_CASES + WARP_CASES + WARP_CASES_FP16 + WARP_CASES_N4)
def test_kernel_source_on_the_host_matches_the_reference_vectors(mirror, name):
    g = load_golden(name)
    f = mirror_run(mirror, g)
    assert f.dtype == g["f_final"].astype(O.policy_dtypes(g["policy"])[1]).dtype
    assert rel_err(f, g["f_final"]) <= RTOL[g["policy"]], rel_err(f, g["f_final"])


@pytest.mark.parametrize("v", [2, 4])
@pytest.mark.parametrize("name", ["sphere_d3q27_kbc_fp32", "warp_tunnel_d3q19_bgk_zouhe_pressure", "cavity_d3q19_bgk_fp32fp16", "channel2d_d2q9_bgk_outflow_fp32"])
def test_cells_per_thread_variants_are_identical_to_one_cell_per_thread(mirror, name, v):
    g = load_golden(name)
    if g["shape"][-1] % v:
        pytest.skip("nz not divisible")
    assert np.array_equal(mirror_run(mirror, g, steps=8, v=v), mirror_run(mirror, g, steps=8, v=1))


BGK_F32 = [n for n in STEP_CASES + LATE_CASES + WARP_CASES + WARP_CASES_FP16 if "bgk" in n]  # every policy, fp64 compute included


@pytest.mark.parametrize("name", BGK_F32)
def test_bgk_source_is_bit_identical_to_the_reference_kernel(mirror, name):
    """The BGK chain of the kernel source (macroscopic -> equilibrium -> BGK, every BC functional) takes one IEEE rounding per operation in
    the reference's operation order (csrc/lbm_math.cuh "ROUNDINGS"; the library is built with -fmad=false), so fp32-compute runs — fp32
    or fp16 storage — reproduce the C restatement of the reference's fused Warp kernel BIT FOR BIT, on the scalar path and on the
    half2-state pair path (whose packed arithmetic and shared-reciprocal division round identically)."""
    from common import c_oracle_run

    g = load_golden(name)
    ref, _, _ = c_oracle_run(g)
    f = mirror_run(mirror, g)
    assert np.array_equal(f, ref), f"scalar path: rel err {rel_err(f, ref):.3e}, {int((f != ref).sum())} values differ"
    if g["policy"] == "FP32FP16" and g["shape"][-1] % 2 == 0 and len(g["shape"]) == 3:
        f2 = mirror_run(mirror, g, v=202)
        assert np.array_equal(f2, ref), f"half2-state path: rel err {rel_err(f2, ref):.3e}, {int((f2 != ref).sum())} values differ"


@pytest.mark.parametrize("cells", [512, 1024])
@pytest.mark.parametrize("lattice,shape,walls", [("D3Q19", (3, 4, 512), True), ("D3Q19", (4, 8, 128), True), ("D3Q27", (3, 16, 64), True), ("D3Q19", (2, 64, 8), False),
                                                 ("D3Q27", (5, 2, 256), False), ("D3Q19", (1, 32, 16), True), ("D3Q19", (2, 128, 8), False)])  # fmt: skip
def test_tile_kernel_logic_is_bit_identical_to_the_reference_kernel(mirror, lattice, shape, walls, cells):
    """cells_per_thread 402 (csrc/step_tile.cuh): copy plan (y wrap, x wrap), z rotation inside the stage rows, boundary handling in the
    output stage — executed on the host with the bulk copies as memcpy — against the C oracle, bit for bit."""
    from common import c_oracle_run

    if shape[1] % (cells // shape[2]):
        pytest.skip("the plane is not a whole number of tiles of this size (the library picks the other size / the direct kernel)")
    g = tile_case(lattice, shape, 6, 11, walls)
    ref, _, _ = c_oracle_run(g)
    f = mirror_run(mirror, g, v=402 if cells == 1024 else 404)
    assert np.array_equal(f, ref), f"{int((f != ref).sum())} of {f.size} values differ, rel err {rel_err(f, ref):.3e}"


@pytest.mark.parametrize("policy", ["FP32FP32", "FP64FP32"])
@pytest.mark.parametrize("lattice,shape,walls", [("D3Q19", (3, 4, 512), True), ("D3Q19", (4, 8, 128), True), ("D3Q27", (3, 16, 64), True), ("D3Q19", (2, 64, 16), False),
                                                 ("D3Q27", (5, 2, 256), False), ("D3Q19", (1, 32, 16), True)])  # fmt: skip
def test_scalar_tile_kernel_logic_is_bit_identical_to_the_reference_kernel(mirror, lattice, shape, walls, policy):
    """cells_per_thread 501 (csrc/step_tile.cuh, step_tile1_kernel): the copy plan (tile_runs: y wrap, x wrap, one-row tiles), the shifted
    reads out of the stage rows and the per-cell code behind them — executed on the host with the bulk copies as memcpy — against the C
    oracle, bit for bit.  (The D3Q19 instantiation of this kernel is where ptxas once merged two arrays of the array-based plan.)"""
    from common import c_oracle_run

    g = tile_case(lattice, shape, 6, 13, walls, policy)
    ref, _, _ = c_oracle_run(g)
    f = mirror_run(mirror, g, v=501)
    assert np.array_equal(f, ref), f"{int((f != ref).sum())} of {f.size} values differ, rel err {rel_err(f, ref):.3e}"


@pytest.mark.parametrize("collision,force", [("BGK", (1e-5, 0.0, -2e-5)), ("SmagorinskyLESBGK", None), ("SmagorinskyLESBGK", (1e-5, 0.0, -2e-5))])
def test_scalar_tile_kernel_logic_with_the_extended_collision_operators(mirror, collision, force):
    """The extended operators on D3Q19 FP32FP32 through the scalar tile kernel's per-thread code: same bits as the direct-load path."""
    from common import c_oracle_run

    g = tile_case("D3Q19", (3, 8, 64), 6, 19, True, "FP32FP32", collision, force)
    ref, _, _ = c_oracle_run(g)
    direct = mirror_run(mirror, g)
    assert rel_err(direct, ref) <= RTOL["FP32FP32"]
    assert np.array_equal(mirror_run(mirror, g, v=501), direct)


def test_scalar_tile_kernel_eligibility_rule(mirror):
    """tile1_eligible (shared by the library's dispatch and the host mirror): nz | 512, nz % 16 == 0, ny a whole number of tiles."""
    for shape in ((4, 16, 16), (2, 24, 32), (2, 64, 8)):  # 32-row tile in a 16-row plane; ny % 16 != 0; nz % 16 != 0
        g = tile_case("D3Q19", shape, 1, 3, False, "FP32FP32")
        with pytest.raises(AssertionError, match="not eligible"):
            mirror_run(mirror, g, v=501)


def test_tile_kernel_logic_with_every_boundary_kind(mirror):
    """Regularized inlet, ExtrapolationOutflow outlet, Halfway body, 255 cells: the scalar boundary routine inside the tile path."""
    from common import c_oracle_run

    for name in ("warp_sphere_d3q19_bgk_fp32fp16", "warp_tunnel_d3q19_bgk_zouhe_pressure_fp32fp16"):
        g0 = load_golden(name)
        nx, ny, nz = g0["shape"]
        # same boundary set on a grid both tile sizes fit (nz = 16, ny = 64): re-derive the index lists
        shape = (nx, 64, 16)
        lat = O.Lattice("D3Q19")
        box, bne = O.bounding_box_indices(shape), O.bounding_box_indices(shape, remove_edges=True)
        walls = np.unique(np.concatenate([box[k] for k in ("bottom", "top", "front", "back")], axis=1), axis=-1)
        X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
        body = np.array(np.where((X - shape[0] // 4) ** 2 + (Y - shape[1] // 2) ** 2 + (Z - shape[2] // 2) ** 2 < 9))
        g = dict(g0)
        g.update(shape=shape, steps=8, f_init=O.initialize_eq(shape, lat, "FP32FP16"))
        bcs = []
        for b in g0["bcs"]:
            b = dict(b)
            if b["kind"] == "fullway":
                b["indices"] = walls
            elif b["kind"] in ("regularized", "zouhe") and b["bc_type"] == "velocity":
                b["indices"] = bne["left"]
                pv = np.zeros((3, shape[1], shape[2]), np.float16)
                pv[0] = 0.03
                b["prescribed"] = pv
            elif b["kind"] == "halfway":
                b["indices"] = body
            else:
                b["indices"] = bne["right"]
            bcs.append(b)
        g["bcs"] = bcs
        g["solid255"] = np.array([[shape[0] // 4], [shape[1] // 2], [shape[2] // 2]])  # one solid cell inside the body
        ref, _, _ = c_oracle_run(g)
        for v in (402, 404):  # 1024- and 512-cell tiles
            f = mirror_run(mirror, g, v=v)
            assert np.array_equal(f, ref), f"{name} v={v}: {int((f != ref).sum())} of {f.size} values differ, rel err {rel_err(f, ref):.3e}"
        # the same set-up with fp32 storage through the scalar tile kernel's per-thread code (cells_per_thread 501)
        for policy in ("FP32FP32", "FP64FP32"):
            g1 = dict(g)
            g1.update(policy=policy, f_init=O.initialize_eq(shape, lat, policy))
            g1["bcs"] = [dict(b, prescribed=np.asarray(b["prescribed"], dtype=np.float32)) if "prescribed" in b and np.ndim(b["prescribed"]) > 0 else b for b in bcs]
            ref1, _, _ = c_oracle_run(g1)
            f1 = mirror_run(mirror, g1, v=501)
            assert np.array_equal(f1, ref1), f"{name} {policy} v=501: {int((f1 != ref1).sum())} of {f1.size} values differ, rel err {rel_err(f1, ref1):.3e}"


@pytest.mark.parametrize("name", ["cavity_d2q9_bgk_fp32", "cavity_d2q9_kbc_fp32", "channel2d_d2q9_bgk_outflow_fp32", "channel2d_d2q9_bgk_zouhe_pressure_fp32", "warp_channel2d_d2q9_bgk_regpressure"])
def test_2d_slab_axis_order_gives_the_same_bits(mirror, name):
    """D2Q9X (csrc/lattice.cuh): the axis order 2-D x-slab runs use — physical x on the kernel's ghost-plane axis, extents (nx, 1, ny) —
    against the ordinary 2-D order (1, nx, ny): identical populations, BC normals and outflow neighbour reads included."""
    g = load_golden(name)
    assert np.array_equal(mirror_run(mirror, g, slab_axes=True), mirror_run(mirror, g))


@pytest.mark.parametrize("name", [n for n in STEP_CASES + LATE_CASES + WARP_CASES if "kbc" in n and "fp16" not in n])
def test_literal_kbc_with_the_references_roundings_is_bit_identical(mirror, name):
    """cells_per_thread = 300 (kExactKbc): the literal KBC formulation with IEEE divisions and nothing fused reproduces the C restatement
    of the reference kernel bit for bit — the fast forms (lean default, reciprocal-based literal) are held to the 1e-5 tolerance."""
    from common import c_oracle_run

    g = load_golden(name)
    ref, _, _ = c_oracle_run(g)
    f = mirror_run(mirror, g, exact_kbc=True)
    assert np.array_equal(f, ref), f"{int((f != ref).sum())} of {f.size} values differ, rel err {rel_err(f, ref):.3e}"
