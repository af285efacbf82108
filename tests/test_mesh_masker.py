"""MeshBoundaryMasker on the CPU: the oracle against the reference's own kernel, the library's per-triangle
device code (compiled for the host) against the oracle, and the published overlap test against an independent
separating-axis implementation and against geometric properties of a closed surface."""

import ctypes as C
import os
import shutil
import subprocess
from collections import deque

import numpy as np
import pytest

from common import GOLDEN_DIR, unpack_bits
from oracle import lbm_numpy as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "xlb_b200", "csrc")
SRC = os.path.join(HERE, "host_math", "mirror_mesh.cu")
OUT = os.path.join(HERE, "host_math", "_build", "libmirror_mesh.so")
MESH_CASES = [f"warp_mesh_{body}_{lat}" for body in ("tetrahedron", "box", "octahedron") for lat in ("d3q19", "d3q27")]


def load_mesh_case(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", MESH_CASES)
def test_oracle_reproduces_the_reference_mesh_masker_bit_for_bit(name):
    """Masks produced by the reference's MeshBoundaryMasker kernel (interpreted, tests/golden/make_golden_warp.py): the oracle's
    literal restatement (edge_test="reference") gives the same bc_mask and missing_mask."""
    g = load_mesh_case(name)
    lat = O.Lattice(str(g["lattice"]))
    shape = tuple(int(s) for s in g["shape"])
    bm, mm = O.build_masks_mesh(g["vertices"], int(g["bc_id"]), np.zeros((1,) + shape, np.uint8), np.zeros((lat.q,) + shape, bool), lat, edge_test="reference")
    assert np.array_equal(bm, g["bc_mask"])
    assert np.array_equal(mm, unpack_bits(g["missing_bits"], lat.q))
    solid = np.argwhere(bm[0] == 255)  # the degenerate edge functions of the reference: solid voxels hug the domain diagonal
    assert np.abs(solid[:, 0] - solid[:, 1]).max() <= 1 and np.abs(solid[:, 1] - solid[:, 2]).max() <= 1 and np.abs(solid[:, 0] - solid[:, 2]).max() <= 1


@pytest.fixture(scope="module")
def mirror():
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    deps = [SRC, os.path.join(CSRC, "mesh_math.cuh"), os.path.join(CSRC, "common.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["nvcc", "-std=c++17", "-O1", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        assert proc.returncode == 0, proc.stderr[-3000:]
    lib = C.CDLL(OUT)
    lib.mirror_mesh_solid.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def mirror_solid(lib, vertices, shape, edge_test):
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    dims = np.array(shape, dtype=np.int32)
    solid = np.zeros(tuple(s + 2 for s in shape), dtype=np.uint8)
    assert lib.mirror_mesh_solid(v.ctypes.data, v.shape[0] // 3, dims.ctypes.data, edge_test, solid.ctypes.data) == 0
    return solid.astype(bool)


def random_soup(seed, n_tri, shape):
    rng = np.random.default_rng(seed)
    centre = rng.uniform(2.0, np.array(shape) - 3.0, size=(n_tri, 1, 3))
    return (centre + rng.uniform(-1.9, 1.9, size=(n_tri, 3, 3))).reshape(-1, 3)


@pytest.mark.parametrize("edge_test", ["schwarz_seidel", "reference"])
def test_device_triangle_code_on_the_host_equals_the_oracle(mirror, edge_test):
    shape = (12, 11, 10)
    soups = [load_mesh_case(n)["vertices"] for n in MESH_CASES[::2]] + [random_soup(s, 60, shape) for s in range(3)]
    for verts in soups:
        want = O.mesh_solid_voxels(verts, shape, edge_test)
        got = mirror_solid(mirror, verts, shape, 0 if edge_test == "schwarz_seidel" else 1)
        assert np.array_equal(got, want)
    assert O.mesh_solid_voxels(soups[0], shape, "schwarz_seidel").sum() > 5 * O.mesh_solid_voxels(soups[0], shape, "reference").sum()


def test_slab_voxelisation_equals_the_slice_of_the_global_one(mirror):
    """x-slabs: every rank voxelises the mesh shifted into its local coordinates, with one cell of halo — the device code's
    result for a slab must be the corresponding slice of the global (padded) solid volume."""
    shape = (12, 11, 10)
    verts = load_mesh_case("warp_mesh_box_d3q27")["vertices"]
    whole = O.mesh_solid_voxels(verts, shape)
    for n_slabs in (2, 3, 4):
        h = shape[0] // n_slabs
        for r in range(n_slabs):
            local = mirror_solid(mirror, verts - np.array([r * h, 0.0, 0.0]), (h,) + shape[1:], 0)
            assert np.array_equal(local, whole[r * h : (r + 1) * h + 2])


# ---- independent check of the published test: separating axes (Akenine-Moeller), float64 ---------------------------------


def sat_overlap(tri, low):
    """Triangle vs unit boxes [low, low+1]: 13 separating axes.  tri (3,3), low (..., 3) -> bool (...)."""
    c = low + 0.5
    v = tri[None, :, :] - c.reshape(-1, 1, 3)  # (m, 3 vertices, 3)
    h = 0.5
    out = np.ones(v.shape[0], dtype=bool)
    e = [tri[1] - tri[0], tri[2] - tri[1], tri[0] - tri[2]]
    axes = [np.eye(3)[k] for k in range(3)] + [np.cross(e[0], e[1])] + [np.cross(np.eye(3)[k], e[i]) for k in range(3) for i in range(3)]
    for a in axes:
        if not np.any(a):
            continue
        p = v @ a
        r = h * np.abs(a).sum()
        out &= ~((p.min(axis=1) > r) | (p.max(axis=1) < -r))
    return out.reshape(low.shape[:-1])


@pytest.mark.parametrize("seed", range(4))
def test_published_overlap_test_agrees_with_separating_axes(seed):
    """Schwarz-Seidel is a reformulation of the exact triangle / box overlap; on generic positions (no touching) the float32
    oracle must agree with a float64 separating-axis test voxel by voxel."""
    shape = (10, 9, 8)
    tris = random_soup(100 + seed, 25, shape).reshape(-1, 3, 3)
    I, J, K = np.meshgrid(*[np.arange(-1, s + 1) for s in shape], indexing="ij")
    low = np.stack([I, J, K], axis=-1).astype(np.float64)
    for tri in tris:
        t = O._tri_setup(*tri, "schwarz_seidel")
        ss = O._tri_box_overlap(t, low.astype(np.float32))
        sat = sat_overlap(tri.astype(np.float32).astype(np.float64), low)
        assert np.array_equal(ss, sat)


@pytest.mark.parametrize("body", ["tetrahedron", "box", "octahedron"])
def test_voxelised_surface_is_conservative_and_watertight(body):
    g = load_mesh_case(f"warp_mesh_{body}_d3q27")
    shape = tuple(int(s) for s in g["shape"])
    verts = g["vertices"]
    solid = O.mesh_solid_voxels(verts, shape)[1:-1, 1:-1, 1:-1]
    # conservative: the voxel of every point sampled on the surface is solid
    rng = np.random.default_rng(0)
    for tri in verts.reshape(-1, 3, 3):
        w = rng.dirichlet(np.ones(3), size=400)
        pts = np.floor(w @ tri).astype(int)
        assert solid[pts[:, 0], pts[:, 1], pts[:, 2]].all()
    # watertight for the 27-neighbourhood's face connectivity: a flood fill of the non-solid cells from a domain corner never
    # reaches the body's centroid cell
    seen = np.zeros(shape, bool)
    queue = deque([(0, 0, 0)])
    seen[0, 0, 0] = True
    while queue:
        i, j, k = queue.popleft()
        for di, dj, dk in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
            a, b, c = i + di, j + dj, k + dk
            if 0 <= a < shape[0] and 0 <= b < shape[1] and 0 <= c < shape[2] and not seen[a, b, c] and not solid[a, b, c]:
                seen[a, b, c] = True
                queue.append((a, b, c))
    centroid = tuple(np.floor(verts.mean(axis=0)).astype(int))
    assert not solid[centroid] and not seen[centroid]
    # mask semantics (mesh_boundary_masker.py:170-188): boundary cells are exactly the non-solid cells with a solid neighbour
    lat = O.Lattice("D3Q27")
    bm, mm = O.build_masks_mesh(verts, 7, np.zeros((1,) + shape, np.uint8), np.zeros((lat.q,) + shape, bool), lat)
    assert np.array_equal(bm[0] == 255, solid)
    pad = np.pad(solid, 1)
    near = np.zeros(shape, bool)
    for l in range(1, lat.q):
        c = lat.c[:, l]
        nb = pad[1 + c[0] : 1 + c[0] + shape[0], 1 + c[1] : 1 + c[1] + shape[1], 1 + c[2] : 1 + c[2] + shape[2]]
        near |= nb
        assert np.array_equal(mm[lat.opp[l]], nb & ~solid)
    assert np.array_equal(bm[0] == 7, near & ~solid)


# ---- the CUDA voxeliser (xlbn_mask_mesh through MeshBoundaryMasker) ---------------------------------------------------------------


def _gpu_mesh_env(lattice):
    import xlb_b200 as xlb
    from xlb_b200.compute_backend import ComputeBackend

    pp, be = xlb.PrecisionPolicy.FP32FP32, ComputeBackend.WARP
    vs = getattr(xlb.velocity_set, lattice)(pp, be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    return xlb, vs, pp, be


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["reference", "schwarz_seidel"])
@pytest.mark.parametrize("name", MESH_CASES)
def test_cuda_mesh_masker_equals_the_oracle_and_the_reference_masks(name, mode):
    """Both edge tests vs the oracle; the literal one also vs the masks the reference's own kernel produced."""
    from xlb_b200.grid import grid_factory
    from xlb_b200.operator.boundary_condition import HalfwayBounceBackBC
    from xlb_b200.operator.boundary_masker import MeshBoundaryMasker

    g = load_mesh_case(name)
    lattice, shape = str(g["lattice"]), tuple(int(s) for s in g["shape"])
    xlb, vs, pp, be = _gpu_mesh_env(lattice)
    grid = grid_factory(shape)
    lat = O.Lattice(lattice)
    bc = HalfwayBounceBackBC(mesh_vertices=g["vertices"].copy())
    bc.id = int(g["bc_id"])
    bc_mask = grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    missing = grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    bc_mask, missing = MeshBoundaryMasker(vs, pp, be, edge_test=mode)(bc, bc_mask, missing)
    bm, mm = O.build_masks_mesh(g["vertices"], bc.id, np.zeros((1,) + shape, np.uint8), np.zeros((lat.q,) + shape, bool), lat, edge_test=mode)
    assert np.array_equal(bc_mask.numpy(), bm) and np.array_equal(missing.numpy(), mm)
    if mode == "reference":
        assert np.array_equal(bc_mask.numpy(), g["bc_mask"]) and np.array_equal(missing.numpy(), unpack_bits(g["missing_bits"], lat.q))


@pytest.mark.gpu
def test_cuda_wind_tunnel_with_a_mesh_body_against_the_c_oracle():
    """examples/cfd/windtunnel_3d.py:66-96 in small: Fullway walls, Regularized inlet, ExtrapolationOutflow, Halfway mesh body, D3Q27 KBC;
    masks bit-exact vs the oracle voxeliser, 20 steps vs the C oracle on the same masks."""
    from common import rel_err
    from oracle import lbm_c
    from xlb_b200.grid import grid_factory
    from xlb_b200.operator.boundary_condition import ExtrapolationOutflowBC, FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    g = load_mesh_case("warp_mesh_octahedron_d3q27")
    shape = (24, 11, 10)
    xlb, vs, pp, be = _gpu_mesh_env("D3Q27")
    grid = grid_factory(shape)
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    verts = g["vertices"] + np.array([2.0, 0.0, 0.0])
    bcs = [FullwayBounceBackBC(indices=walls), RegularizedBC("velocity", prescribed_value=(0.03, 0.0, 0.0), indices=bne["left"]),
           ExtrapolationOutflowBC(indices=bne["right"]), HalfwayBounceBackBC(mesh_vertices=verts.copy())]  # fmt: skip
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="KBC")
    f_0, f_1, bc_mask, missing = stepper.prepare_fields()
    lat = O.Lattice("D3Q27")
    obcs = [O.BC("fullway", bcs[0].id, np.array(walls)), O.BC("regularized", bcs[1].id, np.array(bne["left"]), bc_type="velocity", prescribed=np.array([0.03, 0.0, 0.0])),
            O.BC("outflow", bcs[2].id, np.array(bne["right"])), O.BC("halfway", bcs[3].id, np.zeros((3, 0), np.int64))]  # fmt: skip
    bm, mm = O.build_masks(obcs[:3], shape, lat, flavor="warp")
    bm, mm = O.build_masks_mesh(verts, bcs[3].id, bm, mm, lat)
    assert np.array_equal(bc_mask.numpy(), bm) and np.array_equal(missing.numpy(), mm)
    for i in range(20):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, 1.6, i)
        f_0, f_1 = f_1, f_0
    err = rel_err(f_0.numpy(), lbm_c.run(O.initialize_eq(shape, lat), bm, mm, obcs, 1.6, lat, 20, "FP32FP32", "KBC"))
    assert err <= 1e-5, err
