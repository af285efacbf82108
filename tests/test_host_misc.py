"""CPU-only checks of the pieces around the hot path: the C header is valid C, the reference arm of bench.py prints the
contract's JSON line, the `warp` / `jax` stand-ins evaluate reference-style inlet profiles, example scripts compile."""

import json
import os
import py_compile
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: the header must compile as C (no C++ / torch types)."""
    src = '#include "xlb_b200.h"\nint main(void) { xlbn_domain d = {1, 1, 1, 0, 1}; (void)d; return XLBN_VERSION == 100 ? 0 : 1; }\n'
    proc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"], input=src, text=True, capture_output=True)
    assert proc.returncode == 0, proc.stderr


def test_reference_arm_prints_contract_line():
    from oracle import lbm_c

    if not lbm_c.available():
        pytest.skip("oracle/liblbm_ref.so not built")
    proc = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1", "--n", "64"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["steps"] == 2 and "cavity D3Q19 BGK 64^3 per GPU FP32FP32" in line["config"]["workload"] and line["scaling"] == "weak"
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["unit"] == "MLUPS" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def _bench_module():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_names_the_kernel_the_library_picks():
    """roofline.kernel mirrors the dispatch rules of csrc/step_kernel.cuh (launch_step / launch_step_base)."""
    import argparse

    b = _bench_module()

    def label(shape=(512, 512, 512), **kw):
        return b.kernel_label(argparse.Namespace(**{**dict(lattice="D3Q19", collision="BGK", policy="FP32FP32", force=0.0, cells_per_thread=0), **kw}), shape)

    assert "step_tile1_kernel" in label()
    assert "step_tile1_kernel" in label(policy="FP64FP32")
    assert "step_tile1_kernel" in label(collision="SmagorinskyLESBGK")
    assert "step_kernel (direct" in label(force=1e-5)  # forced operators: on request only
    assert "step_tile1_kernel" in label(force=1e-5, cells_per_thread=501)
    assert "step_kernel (direct" in label(cells_per_thread=1)
    assert "step_kernel (direct" in label(lattice="D3Q27")
    assert "step_kernel (direct" in label(lattice="D3Q27", collision="KBC")
    assert "step_tile_kernel" in label(policy="FP32FP16") and "step_tile_kernel" in label(policy="FP32FP16", lattice="D3Q27")
    assert "step_kernel (direct" in label(shape=(16, 16, 16))  # a 512-cell tile would be 32 rows of a 16-row plane
    assert "step_kernel (direct" in label(shape=(64, 24, 32))  # ny is not a whole number of tiles


def test_clock_sampler_falls_back_without_nvml():
    """No GPU here: NVML cannot initialise, bench.py must fall back to the nvidia-smi sampler and still return a clocks object."""
    import torch

    b = _bench_module()
    s = b.clock_sampler(torch, 0)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the NVML sampler is used")
    assert type(s).__name__ == "ClockSampler"
    s.start()
    out = s.stop(0.0, 1.0)
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert proc.returncode == 0 and proc.stdout.strip() == ""


def test_warp_and_jax_standins_are_installed_only_when_missing():
    import xlb  # noqa: F401  (installs the stand-ins)
    import warp as wp
    import jax.numpy as jnp

    assert getattr(wp, "__standin__", False), "a real warp is installed; the stand-in must not shadow it"
    H = 9.0

    @wp.func
    def profile(index: wp.vec3i):
        y = wp.float32(index[1])
        return wp.vec(0.04 * wp.max(0.0, 1.0 - (2.0 * (y - H / 2.0) / H) ** 2.0), length=1)

    from xlb_b200.operator.boundary_condition.bc_zouhe import _evaluate_index_profile

    cells = np.array([[0, 0, 0], [0, 4, 9], [1, 2, 3]])
    vec = _evaluate_index_profile(profile, cells)
    one_by_one = np.array([profile((0, int(y), 0))[0] for y in cells[1]])
    assert vec.shape == (3,) and np.allclose(vec, one_by_one)

    def scalar_only(index):  # not array-aware: the per-cell fallback must kick in
        return [0.01 * float(int(index[1]) % 3)]

    assert np.allclose(_evaluate_index_profile(scalar_only, cells), [0.0, 0.01, 0.0])
    u = torch.tensor([[3.0, 0.0], [4.0, 1.0]])
    assert isinstance(u, jnp.ndarray) and torch.allclose(jnp.sqrt(u[0] ** 2 + u[1] ** 2), torch.tensor([5.0, 1.0]))
    assert np.allclose(jnp.maximum(0.0, np.array([-1.0, 2.0])), [0.0, 2.0])


@pytest.mark.parametrize("script", ["examples/cavity_mlups.py", "examples/sphere_kbc.py", "examples/cavity_2d.py", "examples/windtunnel_mesh.py", "examples/turbulent_channel.py", "bench.py", "__graft_entry__.py", "scripts/mgpu_check.py"])
def test_scripts_compile(script):
    py_compile.compile(os.path.join(ROOT, script), doraise=True)


def test_stepper_rejects_out_of_scope_options_loudly():
    import xlb_b200 as xlb
    from xlb_b200.compute_backend import ComputeBackend
    from xlb_b200.grid import grid_factory
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    pp = xlb.PrecisionPolicy.FP32FP32
    xlb.init(velocity_set=xlb.velocity_set.D3Q19(pp, ComputeBackend.WARP), default_backend=ComputeBackend.WARP, default_precision_policy=pp)
    g = grid_factory((4, 4, 4), device="cpu")
    for kw in (dict(collision_type="MRT"), dict(streaming_scheme="push"), dict(collision_type="KBC")):
        with pytest.raises(NotImplementedError):
            IncompressibleNavierStokesStepper(grid=g, boundary_conditions=[], **kw)
    with pytest.raises(AssertionError):  # forced_collision.py:30 "Check the dimensions of the input force!"
        IncompressibleNavierStokesStepper(grid=g, boundary_conditions=[], force_vector=np.zeros(2))
    with pytest.raises(AssertionError):  # forced_collision.py:27: only "exact_difference" exists
        IncompressibleNavierStokesStepper(grid=g, boundary_conditions=[], force_vector=np.zeros(3), forcing_scheme="guo")


def test_stepper_collision_options_mirror_the_reference_ctor():
    """nse_stepper.py:38-46: collision_type in {BGK, KBC, SmagorinskyLESBGK}, optionally wrapped by ForcedCollision."""
    import xlb_b200 as xlb
    from xlb_b200 import native
    from xlb_b200.compute_backend import ComputeBackend
    from xlb_b200.grid import grid_factory
    from xlb_b200.operator.collision import BGK, KBC, ForcedCollision, SmagorinskyLESBGK
    from xlb_b200.operator.force import ExactDifference
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    pp = xlb.PrecisionPolicy.FP32FP32
    xlb.init(velocity_set=xlb.velocity_set.D3Q27(pp, ComputeBackend.WARP), default_backend=ComputeBackend.WARP, default_precision_policy=pp)
    g = grid_factory((4, 4, 4), device="cpu")
    s = IncompressibleNavierStokesStepper(grid=g, collision_type="SmagorinskyLESBGK")
    assert isinstance(s.collision, SmagorinskyLESBGK) and s.collision.smagorinsky_coef == 0.17 and s.collision.native_collision == native.SMAGORINSKY_LES_BGK
    s = IncompressibleNavierStokesStepper(grid=g, collision_type="KBC", force_vector=np.array([1e-5, 0.0, 0.0]))
    assert isinstance(s.collision, ForcedCollision) and isinstance(s.collision.collision_operator, KBC)
    assert isinstance(s.collision.forcing_operator, ExactDifference)
    assert s.collision.native_collision == native.KBC | native.COLLISION_FORCED and list(s.collision.native_force) == [1e-5, 0.0, 0.0]
    s = IncompressibleNavierStokesStepper(grid=g, collision_type="SmagorinskyLESBGK", force_vector=np.array([0.0, 2e-5, 0.0]))
    assert s.collision.native_collision == native.SMAGORINSKY_LES_BGK | native.COLLISION_FORCED and s.collision.native_smagorinsky == 0.17
    assert isinstance(IncompressibleNavierStokesStepper(grid=g).collision, BGK)
    xlb.init(velocity_set=xlb.velocity_set.D2Q9(pp, ComputeBackend.WARP), default_backend=ComputeBackend.WARP, default_precision_policy=pp)
    with pytest.raises(NotImplementedError):  # the reference functional reads c[2, l]: 3-D only
        SmagorinskyLESBGK()


def test_read_stl_binary_and_ascii(tmp_path):
    from xlb_b200.utils import read_stl

    tris = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1], [1, 0.5, 1], [0.25, 1, 1]]], dtype=np.float32)
    rec = np.zeros(2, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]))
    rec["v"] = tris
    binary = tmp_path / "b.stl"
    binary.write_bytes(b"binary stl".ljust(80, b" ") + np.uint32(2).tobytes() + rec.tobytes())
    ascii_ = tmp_path / "a.stl"
    body = "".join("facet normal 0 0 1\n outer loop\n" + "".join(f"  vertex {x} {y} {z}\n" for x, y, z in t) + " endloop\nendfacet\n" for t in tris)
    ascii_.write_text("solid s\n" + body + "endsolid s\n")
    for path in (binary, ascii_):
        v = read_stl(str(path))
        assert v.shape == (6, 3) and v.dtype == np.float64 and np.allclose(v, tris.reshape(-1, 3))
    bad = tmp_path / "c.stl"
    bad.write_text("hello")
    with pytest.raises(ValueError):
        read_stl(str(bad))


def test_mesh_boundary_masker_argument_checks_mirror_the_reference():
    """mesh_boundary_masker.py:27-28 (2-D), L199-216 (vertices / indices / (N, 3) / mesh inside the domain) — checked before
    anything touches the device, in the reference's order; and there is no CPU fallback behind them."""
    import torch

    import xlb_b200 as xlb
    from xlb_b200.compute_backend import ComputeBackend
    from xlb_b200.operator.boundary_condition import HalfwayBounceBackBC
    from xlb_b200.operator.boundary_masker import MeshBoundaryMasker

    pp, be = xlb.PrecisionPolicy.FP32FP32, ComputeBackend.WARP
    xlb.init(velocity_set=xlb.velocity_set.D2Q9(pp, be), default_backend=be, default_precision_policy=pp)
    with pytest.raises(NotImplementedError, match="not implemented in 2D"):
        MeshBoundaryMasker()
    vs = xlb.velocity_set.D3Q19(pp, be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    with pytest.raises(ValueError, match="edge_test"):
        MeshBoundaryMasker(edge_test="sat")
    masker = MeshBoundaryMasker()
    bc_mask, missing = torch.zeros(1, 8, 8, 8, dtype=torch.uint8), torch.zeros(19, 8, 8, 8, dtype=torch.bool)
    tri = np.array([[1.0, 1.0, 1.0], [3.0, 1.0, 1.0], [1.0, 3.0, 2.0]])
    with pytest.raises(Exception, match="Please provide the mesh vertices"):
        masker(HalfwayBounceBackBC(indices=[[1], [1], [1]]), bc_mask, missing)
    with pytest.raises(Exception, match=r"reshaped into an array \(N, 3\)"):
        masker(HalfwayBounceBackBC(mesh_vertices=tri[:, :2]), bc_mask, missing)
    with pytest.raises(Exception, match="three consecutive rows per triangle"):
        masker(HalfwayBounceBackBC(mesh_vertices=tri[:2]), bc_mask, missing)
    with pytest.raises(Exception, match="exceed domain dimensions"):
        masker(HalfwayBounceBackBC(mesh_vertices=tri + 6.0), bc_mask, missing)
    bc = HalfwayBounceBackBC(mesh_vertices=tri)
    with pytest.raises(Exception, match="no CPU fallback"):
        masker(bc, bc_mask, missing)
    assert bc.mesh_vertices is not None  # only a successful call consumes the vertices (reference L212-213)
