"""Golden vectors from THE REFERENCE'S WARP BACKEND, executed here cell by cell.

Run in the build container only:   python tests/golden/make_golden_warp.py      (≈ 10 min; pure-Python per-cell execution)

`warp` is not installable here, but Warp kernels are syntactically Python: oracle/refshim's interpretive `warp` stand-in
(`install(interpret_warp=True)`) turns `wp.func` into a by-value call, `wp.kernel` + `wp.launch` into a loop over the launch
grid and `wp.vec / wp.mat` into small numpy value types.  Everything that runs below — the fused step kernel
(xlb/operator/stepper/nse_stepper.py:344-381), Stream / Macroscopic / QuadraticEquilibrium / BGK / KBC /
SmagorinskyLESBGK / ForcedCollision functionals, every boundary-condition functional incl. the aux-data kernels
(boundary_condition.py:119-175), the Warp IndicesBoundaryMasker kernel (indices_boundary_masker.py:111-159) and
MomentumTransfer (momentum_transfer.py:92-176) — is the reference's own source.  These cases pin what the JAX-path vectors
(make_golden.py) cannot: the 255 skip, the scalar prescribed value kept in f_1[0, cell] and its recovery, the per-index
interior flag of the Warp masker, the Warp outflow neighbour read, and SmagorinskyLESBGK (which has no JAX implementation).

File format = make_golden.py's (tests/common.py:load_golden), plus `backend="WARP"`, optional `solid255` (cells whose
bc_mask was set to 255 after prepare_fields, as MeshBoundaryMasker does for solid interiors), optional `smagorinsky`.
FP32FP32, plus one FP64FP64 case (with omega representable in float32: Warp types a Python float handed to a launch as
float32, so any other omega would be rounded there before the kernel widens it); the mixed policies depend on implicit
conversions of Warp that the stand-in does not model.
"""

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

os.chdir("/tmp")  # keep the repo's own `xlb` alias package off the import path
if ROOT in sys.path:
    sys.path.remove(ROOT)
xlb = refshim.import_reference("/root/reference", interpret_warp=True)
sys.path.append(ROOT)

import warp as wp  # noqa: E402  (the interpretive stand-in)
from xlb.compute_backend import ComputeBackend  # noqa: E402
from xlb.precision_policy import PrecisionPolicy  # noqa: E402
from xlb.grid import grid_factory  # noqa: E402
from xlb.operator.stepper import IncompressibleNavierStokesStepper  # noqa: E402
from xlb.operator.macroscopic import Macroscopic  # noqa: E402
from xlb.operator.equilibrium import QuadraticEquilibrium  # noqa: E402
from xlb.operator.boundary_condition import (  # noqa: E402
    DoNothingBC,
    EquilibriumBC,
    ExtrapolationOutflowBC,
    FullwayBounceBackBC,
    HalfwayBounceBackBC,
    RegularizedBC,
    ZouHeBC,
)
from xlb.operator.boundary_condition.boundary_condition_registry import boundary_condition_registry  # noqa: E402

BE = ComputeBackend.WARP
VS = {"D2Q9": xlb.velocity_set.D2Q9, "D3Q19": xlb.velocity_set.D3Q19, "D3Q27": xlb.velocity_set.D3Q27}
POLICY = "FP32FP32"


def init(lattice, policy="FP32FP32"):
    global POLICY
    POLICY = policy
    pp = PrecisionPolicy[POLICY]
    # one case per "process": the Warp stepper looks outflow ids up in the GLOBAL registry (nse_stepper.py:254-259), so ids of an
    # earlier case must not linger
    boundary_condition_registry.next_id = 1
    boundary_condition_registry.bc_to_id.clear()
    boundary_condition_registry.id_to_bc.clear()
    vs = VS[lattice](precision_policy=pp, compute_backend=BE)
    xlb.init(velocity_set=vs, default_backend=BE, default_precision_policy=pp)
    return vs, pp


def pack_bits(missing):
    m = np.asarray(missing).astype(np.uint32)
    return sum(m[l] << np.uint32(l) for l in range(m.shape[0])).astype(np.uint32)


def run_and_save(name, meta, stepper, bcs_meta, steps, omega, d, f_init=None, solid255=None, force_bc=None, inlet=None):
    t0 = time.time()
    f_0, f_1, bc_mask, missing = stepper.prepare_fields()
    if f_init is not None:  # user-supplied initial state, as an `initializer` would return it
        f_0[...] = f_init
        keep = np.asarray(f_1[0]).copy()  # aux data already written by aux_data_init
        f_1[...] = f_init
        for b in stepper.boundary_conditions:
            if getattr(b, "needs_aux_init", False):
                sel = np.asarray(bc_mask[0]) == b.id
                f_1[0][sel] = keep[sel]
    if solid255 is not None:
        bc_mask[0][tuple(solid255)] = 255
    if inlet is not None:  # the per-cell prescribed scalar exactly as aux_data_init stored it (boundary_condition.py:151)
        i, b = inlet
        idx = bcs_meta[i]["indices"]
        idx3 = tuple(idx) if len(idx) == 3 else (idx[0], idx[1], np.zeros_like(idx[0]))
        pv = np.zeros((d,) + tuple(np.asarray(f_0).shape[2:]), np.asarray(f_1).dtype)  # same container as the JAX-path fixtures: [d, ny, nz]
        pv[0][idx3[1:]] = np.asarray(f_1)[0][idx3]
        bcs_meta[i]["prescribed"] = pv if d == 3 else pv[:, :, 0]
    start = np.asarray(f_0).copy()
    for i in range(steps):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, omega, i)
        f_0, f_1 = f_1, f_0
    cdt = wp.float64 if POLICY.startswith("FP64") else wp.float32
    rho = wp.zeros((1,) + f_0.shape[1:], dtype=cdt)
    u = wp.zeros((3,) + f_0.shape[1:], dtype=cdt) if d == 3 else None
    out = dict(meta)
    out.update(steps=steps, omega=omega, n_bc=len(bcs_meta), backend="WARP", policy=POLICY)
    for i, b in enumerate(bcs_meta):
        for k, v in b.items():
            out[f"bc{i}_{k}"] = v
    squeeze = (lambda a: np.asarray(a)[..., 0]) if d == 2 else np.asarray
    out["f_init"] = squeeze(start)
    out["f_final"] = squeeze(f_0)
    out["bc_mask"] = squeeze(bc_mask)
    out["missing_bits"] = pack_bits(squeeze(missing))
    if d == 3:
        rho, u = Macroscopic()(f_0, rho, u)
        out["rho"], out["u"] = np.asarray(rho), np.asarray(u)
    if solid255 is not None:
        out["solid255"] = np.asarray(solid255)
    if force_bc is not None:
        from xlb.operator.force.momentum_transfer import MomentumTransfer

        out["force"] = np.asarray(MomentumTransfer(force_bc)(f_0, f_1, bc_mask, missing))
        out["force_bc"] = np.int64(stepper.boundary_conditions.index(force_bc))
    assert np.isfinite(out["f_final"]).all(), name
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {out['f_final'].shape} x {steps} steps in {time.time() - t0:.0f} s -> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def cavity(name, lattice, n, steps, collision="BGK", omega=1.0, solid_block=None):
    """examples/performance/mlups_3d.py:45-63 / examples/cfd/lid_driven_cavity_2d.py on the WARP backend."""
    vs, pp = init(lattice)
    shape = (n,) * vs.d
    grid = grid_factory(shape)
    box, box_ne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    lid = box_ne["top"]
    names = ["bottom", "left", "right"] + (["front", "back"] if vs.d == 3 else [])
    walls = [sum((box[k][i] for k in names), []) for i in range(vs.d)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    u_lid = (0.02, 0.0, 0.0)[: vs.d]
    bcs = [EquilibriumBC(rho=1.0, u=u_lid, indices=lid), FullwayBounceBackBC(indices=walls)]
    meta = [
        dict(kind="equilibrium", id=bcs[0].id, indices=np.array(lid), rho=1.0, u=np.array(u_lid)),
        dict(kind="fullway", id=bcs[1].id, indices=np.array(walls)),
    ]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=collision)
    solid = None
    if solid_block is not None:
        lo, hi = solid_block
        solid = np.array(np.nonzero(np.ones([hi - lo] * vs.d, bool))) + lo
    run_and_save(name, dict(lattice=lattice, collision=collision, shape=np.array(shape)), stepper, meta, steps, omega, vs.d, solid255=solid)


def tunnel(name, lattice, shape, steps, collision, omega, inlet="regularized", outlet="outflow", constant_inlet=None, policy="FP32FP32"):
    """examples/cfd/flow_past_sphere_3d.py:41-109 on the WARP backend (per-cell wp.func inlet profile, L86-99)."""
    vs, pp = init(lattice, policy)
    d = vs.d
    u_max = 0.04
    grid = grid_factory(shape)
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    inlet_idx, outlet_idx = bne["left"], bne["right"]
    wall_names = ["bottom", "top"] + (["front", "back"] if d == 3 else [])
    walls = [sum((box[k][i] for k in wall_names), []) for i in range(d)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    axes = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    centre = [shape[0] // 5] + [s // 2 for s in shape[1:]]
    r = max(2, shape[1] // 6)
    ind = np.where(sum((a - c) ** 2 for a, c in zip(axes, centre)) < r**2)
    body = [tuple(ind[i]) for i in range(d)]
    H_y = float(shape[1] - 1)
    H_z = float(shape[2] - 1) if d == 3 else 1.0

    @wp.func
    def profile(index: wp.vec3i):
        yc = wp.float32(index[1]) - (H_y / 2.0)
        r2 = (2.0 * yc / H_y) ** 2.0
        if d == 3:
            zc = wp.float32(index[2]) - (H_z / 2.0)
            r2 = r2 + (2.0 * zc / H_z) ** 2.0
        return wp.vec(u_max * wp.max(0.0, 1.0 - r2), length=1)

    bc_walls = FullwayBounceBackBC(indices=walls)
    In = RegularizedBC if inlet == "regularized" else ZouHeBC
    if constant_inlet is None:
        bc_in = In("velocity", profile=profile, indices=inlet_idx)
    else:
        bc_in = In("velocity", prescribed_value=(constant_inlet,) + (0.0,) * (d - 1), indices=inlet_idx)
    if outlet == "outflow":
        bc_out, out_meta = ExtrapolationOutflowBC(indices=outlet_idx), dict(kind="outflow")
    elif outlet == "pressure":  # a 1-tuple: the WARP branch of the ctor indexes the value (bc_zouhe.py:96-98), a bare float does not survive it
        bc_out = ZouHeBC("pressure", prescribed_value=(1.0,), indices=outlet_idx)
        out_meta = dict(kind="zouhe", bc_type="pressure", prescribed=np.float64(1.0))
    elif outlet == "regularized_pressure":
        bc_out = RegularizedBC("pressure", prescribed_value=(1.0,), indices=outlet_idx)
        out_meta = dict(kind="regularized", bc_type="pressure", prescribed=np.float64(1.0))
    else:
        bc_out, out_meta = DoNothingBC(indices=outlet_idx), dict(kind="donothing")
    bc_body = HalfwayBounceBackBC(indices=body)
    bcs = [bc_walls, bc_in, bc_out, bc_body]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=collision)
    out_meta.update(id=bc_out.id, indices=np.array(outlet_idx))
    meta = [
        dict(kind="fullway", id=bc_walls.id, indices=np.array(walls)),
        dict(kind=inlet, id=bc_in.id, indices=np.array(inlet_idx), bc_type="velocity"),  # `prescribed` filled from f_1[0]
        out_meta,
        dict(kind="halfway", id=bc_body.id, indices=np.array(body)),
    ]
    run_and_save(name, dict(lattice=lattice, collision=collision, shape=np.array(shape)), stepper, meta, steps, omega, d,
                 force_bc=bc_body if d == 3 else None, inlet=(1, bc_in))  # fmt: skip


def periodic(name, lattice, shape, steps, collision, omega, force=None):
    """Fully periodic box from a seeded random state; optional body force (ForcedCollision + ExactDifference, Warp functionals)."""
    vs, pp = init(lattice)
    grid = grid_factory(shape)
    rng = np.random.default_rng(0)
    u0 = (1e-2 * rng.standard_normal((vs.d,) + tuple(shape))).astype(np.float32)
    rho0 = (1.0 + 1e-3 * rng.standard_normal((1,) + tuple(shape))).astype(np.float32)
    tail = (1,) if vs.d == 2 else ()  # Warp fields of 2-D runs carry a trailing singleton axis (warp_grid.py:26)
    f_init = wp.zeros((vs.q,) + tuple(shape) + tail, dtype=wp.float32)
    f_init = QuadraticEquilibrium()(wp.array(rho0.reshape(rho0.shape + tail), dtype=wp.float32), wp.array(u0.reshape(u0.shape + tail), dtype=wp.float32), f_init)
    kw = {} if force is None else dict(force_vector=np.asarray(force, dtype=np.float32))
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=[], collision_type=collision, **kw)
    meta = dict(lattice=lattice, collision=collision, shape=np.array(shape), rho0=rho0, u0=u0)
    if force is not None:
        meta["force_vector"] = np.asarray(force, dtype=np.float64)
    if collision == "SmagorinskyLESBGK":
        meta["smagorinsky"] = np.float64(stepper.collision.smagorinsky_coef if force is None else stepper.collision.collision_operator.smagorinsky_coef)
    run_and_save(name, meta, stepper, [], steps, omega, vs.d, f_init=np.asarray(f_init))


def channel(name, shape, steps, omega, force):
    """examples/cfd/turbulent_channel_3d.py:60-135 in small: periodic in x and y, RegularizedBC("velocity", (0, 0, 0)) on the
    two z faces without their edges, KBC + body force (ForcedCollision / ExactDifference), FP64FP64, seeded random start
    through helper.initialize_eq(u=...)."""
    from xlb.helper import initialize_eq

    vs, pp = init("D3Q27", "FP64FP64")
    grid = grid_factory(shape)
    box = grid.bounding_box_indices(remove_edges=True)
    walls = [box["bottom"][i] + box["top"][i] for i in range(vs.d)]
    bcs = [RegularizedBC("velocity", prescribed_value=(0.0, 0.0, 0.0), indices=walls)]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type="KBC", force_vector=np.asarray(force, dtype=np.float64))
    np.random.seed(0)
    u_init = wp.array(1e-2 * np.random.random((vs.d,) + tuple(shape)), dtype=wp.float64)
    f_init = initialize_eq(wp.zeros((vs.q,) + tuple(shape), dtype=wp.float64), grid, vs, pp, BE, u=u_init)
    meta = [dict(kind="regularized", id=bcs[0].id, indices=np.array(walls), bc_type="velocity")]  # `prescribed` filled from f_1[0]
    run_and_save(name, dict(lattice="D3Q27", collision="KBC", shape=np.array(shape), force_vector=np.asarray(force, dtype=np.float64)), stepper, meta,
                 steps, omega, vs.d, f_init=np.asarray(f_init), inlet=(0, bcs[0]))  # fmt: skip


def mesh_shapes():
    """Small closed triangle soups at generic (non-lattice-aligned) positions: a tetrahedron, a rotated box, an octahedron."""
    tet = np.array([[2.21, 1.37, 1.11], [9.63, 3.19, 2.87], [5.02, 9.43, 3.33], [5.57, 4.21, 8.79]])
    tet_f = [(0, 1, 2), (0, 3, 1), (1, 3, 2), (2, 3, 0)]
    a, b = 0.37, 0.21  # rotated box
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    corners = np.array([[x, y, z] for x in (-2.6, 2.6) for y in (-1.9, 1.9) for z in (-1.4, 1.4)]) @ R.T + np.array([6.13, 5.41, 4.87])
    box_f = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
    octa = np.array([[3.3, 0, 0], [-3.3, 0, 0], [0, 2.9, 0], [0, -2.9, 0], [0, 0, 2.7], [0, 0, -2.7]]) + np.array([6.07, 5.53, 4.61])
    octa_f = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    return {"tetrahedron": (tet, tet_f), "box": (corners, box_f), "octahedron": (octa, octa_f)}


def mesh_masks(name, lattice, shape, body):
    """MeshBoundaryMasker.warp_implementation (mesh_boundary_masker.py:193-236) on zeroed masks, through the reference's own
    kernel (triangle / voxel test L60-148) with wp.Mesh / wp.mesh_query_aabb served by the stand-in's brute-force mesh."""
    from xlb.operator.boundary_masker import MeshBoundaryMasker

    vs, pp = init(lattice)
    grid = grid_factory(shape)
    P, faces = mesh_shapes()[body]
    verts = np.array([P[i] for f in faces for i in f], dtype=np.float64)
    bc = HalfwayBounceBackBC(mesh_vertices=verts.copy())
    bc_mask = grid.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    missing = grid.create_field(cardinality=vs.q, dtype=xlb.Precision.BOOL)
    t0 = time.time()
    bc_mask, missing = MeshBoundaryMasker(vs, pp, BE)(bc, bc_mask, missing)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, lattice=lattice, shape=np.array(shape), vertices=verts, bc_id=np.int64(bc.id), bc_mask=np.asarray(bc_mask),
                        missing_bits=pack_bits(missing), backend="WARP")  # fmt: skip
    print(f"{name}: {int((np.asarray(bc_mask) == 255).sum())} solid, {int((np.asarray(bc_mask) == bc.id).sum())} boundary cells in {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    only = sys.argv[1:]

    def want(n):
        return not only or any(o in n for o in only)

    if want("warp_cavity_d3q19_bgk"):
        cavity("warp_cavity_d3q19_bgk", "D3Q19", 10, 20)
    if want("warp_cavity_d3q19_bgk_solid255"):
        cavity("warp_cavity_d3q19_bgk_solid255", "D3Q19", 10, 12, omega=1.4, solid_block=(4, 6))
    if want("warp_cavity_d2q9_kbc"):
        cavity("warp_cavity_d2q9_kbc", "D2Q9", 16, 30, collision="KBC", omega=1.8)
    if want("warp_tunnel_d3q27_kbc_regularized_outflow"):
        tunnel("warp_tunnel_d3q27_kbc_regularized_outflow", "D3Q27", (24, 12, 12), 15, "KBC", 1.6)
    if want("warp_tunnel_d3q19_bgk_zouhe_pressure"):
        tunnel("warp_tunnel_d3q19_bgk_zouhe_pressure", "D3Q19", (20, 10, 10), 15, "BGK", 1.4, inlet="zouhe", outlet="pressure", constant_inlet=0.03)
    if want("warp_tunnel_d3q19_bgk_regularized_donothing"):
        tunnel("warp_tunnel_d3q19_bgk_regularized_donothing", "D3Q19", (18, 10, 10), 12, "BGK", 1.2, outlet="donothing")
    if want("warp_channel2d_d2q9_bgk"):
        tunnel("warp_channel2d_d2q9_bgk", "D2Q9", (30, 14), 30, "BGK", 1.6)
    if want("warp_channel2d_d2q9_bgk_regpressure"):
        tunnel("warp_channel2d_d2q9_bgk_regpressure", "D2Q9", (30, 14), 30, "BGK", 1.5, inlet="zouhe", outlet="regularized_pressure")
    if want("warp_periodic_d3q27_kbc"):
        periodic("warp_periodic_d3q27_kbc", "D3Q27", (8, 6, 6), 10, "KBC", 1.8)
    if want("warp_periodic_d3q19_bgk_forced"):
        periodic("warp_periodic_d3q19_bgk_forced", "D3Q19", (8, 6, 6), 10, "BGK", 1.5, force=(1e-5, 0.0, 0.0))
    if want("warp_periodic_d3q19_smagorinsky"):
        periodic("warp_periodic_d3q19_smagorinsky", "D3Q19", (8, 6, 6), 10, "SmagorinskyLESBGK", 1.9)
    if want("warp_periodic_d3q27_smagorinsky_forced"):
        periodic("warp_periodic_d3q27_smagorinsky_forced", "D3Q27", (8, 6, 6), 8, "SmagorinskyLESBGK", 1.95, force=(2e-5, -1e-5, 0.0))
    if want("warp_periodic_d3q19_smagorinsky_forced"):
        periodic("warp_periodic_d3q19_smagorinsky_forced", "D3Q19", (6, 8, 4), 8, "SmagorinskyLESBGK", 1.9, force=(0.0, 1e-5, 2e-5))
    if want("warp_periodic_d3q27_smagorinsky"):
        periodic("warp_periodic_d3q27_smagorinsky", "D3Q27", (6, 6, 8), 8, "SmagorinskyLESBGK", 1.85)
    if want("warp_periodic_d3q27_bgk_forced"):
        periodic("warp_periodic_d3q27_bgk_forced", "D3Q27", (8, 6, 4), 8, "BGK", 1.6, force=(1e-5, 2e-5, -1e-5))
    if want("warp_periodic_d3q27_kbc_forced"):
        periodic("warp_periodic_d3q27_kbc_forced", "D3Q27", (6, 8, 6), 8, "KBC", 1.8, force=(-1e-5, 0.0, 2e-5))
    if want("warp_periodic_d2q9_bgk_forced"):
        periodic("warp_periodic_d2q9_bgk_forced", "D2Q9", (12, 10), 12, "BGK", 1.5, force=(2e-5, -1e-5))
    if want("warp_channel_d3q27_kbc_forced_fp64"):
        channel("warp_channel_d3q27_kbc_forced_fp64", (10, 8, 9), 12, 1.75, (3e-6, 0.0, 0.0))
    for body in ("tetrahedron", "box", "octahedron"):
        for lattice in ("D3Q19", "D3Q27"):
            n = f"warp_mesh_{body}_{lattice.lower()}"
            if want(n):
                mesh_masks(n, lattice, (12, 11, 10), body)
    if want("warp_periodic_d2q9_kbc_forced"):
        periodic("warp_periodic_d2q9_kbc_forced", "D2Q9", (10, 12), 12, "KBC", 1.8, force=(1e-5, 1e-5))
    # FP32FP16 on the WARP backend: the prescribed inlet value lives in f_1[0, cell] in the STORE dtype (boundary_condition.py:151,
    # bc_zouhe.py:95-98), every load / store carries an explicit compute_dtype(...) / store_dtype(...) cast in the reference source
    if want("warp_sphere_d3q19_bgk_fp32fp16"):
        tunnel("warp_sphere_d3q19_bgk_fp32fp16", "D3Q19", (32, 14, 14), 30, "BGK", 1.5, policy="FP32FP16")
    if want("warp_tunnel_d3q27_kbc_fp32fp16"):
        tunnel("warp_tunnel_d3q27_kbc_fp32fp16", "D3Q27", (24, 12, 12), 20, "KBC", 1.6, policy="FP32FP16")
    if want("warp_tunnel_d3q19_bgk_zouhe_pressure_fp32fp16"):
        tunnel("warp_tunnel_d3q19_bgk_zouhe_pressure_fp32fp16", "D3Q19", (20, 10, 10), 20, "BGK", 1.4, inlet="zouhe", outlet="pressure", constant_inlet=0.03, policy="FP32FP16")
