"""Generate the golden vectors under tests/golden/ by running THE REFERENCE'S OWN PYTHON (Autodesk/XLB, JAX backend).

Run in the build container only:   python tests/golden/make_golden.py
(/root/reference does not exist on the GPU box; the committed .npz files travel instead.)

Real `jax` / `warp` are not installable here, so the reference is imported under oracle/refshim: numpy-backed
stand-ins for the array primitives (`jnp.roll/where/tensordot/.at[].set/...`) with JAX's dtype rules.  Every line of
LBM arithmetic that produces these vectors is the reference's (xlb/operator/**, JAX implementations); nothing from
xlb_b200 or oracle/lbm_numpy.py is involved.

Each case stores: the inputs needed to rebuild it (shape, lattice, policy, collision, omega, steps, BC index lists and
parameters, ids), the reference's masks, and the populations after `steps` steps of the user loop
(examples/performance/mlups_3d.py:77-80) plus rho/u from the reference's Macroscopic.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

os.chdir("/tmp")  # keep the repo's own `xlb` alias package off the import path
if ROOT in sys.path:
    sys.path.remove(ROOT)
xlb = refshim.import_reference("/root/reference")
sys.path.append(ROOT)

import jax.numpy as jnp  # noqa: E402  (the stand-in)
from xlb.compute_backend import ComputeBackend  # noqa: E402
from xlb.precision_policy import PrecisionPolicy  # noqa: E402
from xlb.grid import grid_factory  # noqa: E402
from xlb.operator.stepper import IncompressibleNavierStokesStepper  # noqa: E402
from xlb.operator.macroscopic import Macroscopic  # noqa: E402
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
from xlb.helper import initialize_eq  # noqa: E402

BE = ComputeBackend.JAX
VS = {"D2Q9": xlb.velocity_set.D2Q9, "D3Q19": xlb.velocity_set.D3Q19, "D3Q27": xlb.velocity_set.D3Q27}


def init(lattice, policy):
    pp = PrecisionPolicy[policy]
    refshim.set_x64(policy.startswith("FP64"))
    boundary_condition_registry.next_id = 1
    vs = VS[lattice](precision_policy=pp, compute_backend=BE)
    xlb.init(velocity_set=vs, default_backend=BE, default_precision_policy=pp)
    return vs, pp


def pack_bits(missing):
    m = np.asarray(missing).astype(np.uint32)
    return sum(m[l] << np.uint32(l) for l in range(m.shape[0])).astype(np.uint32)


def run_and_save(name, meta, stepper, bcs_meta, steps, omega, initializer=None, force_bc=None):
    f_0, f_1, bc_mask, missing = stepper.prepare_fields(initializer=initializer)
    f_init = np.asarray(f_0).copy()
    for i in range(steps):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing, omega, i)
        f_0, f_1 = f_1, f_0
    macro = Macroscopic()
    rho, u = macro(jnp.array(f_0, dtype=xlb.DefaultConfig.default_precision_policy.compute_precision.jax_dtype))
    out = dict(meta)
    out.update(steps=steps, omega=omega, n_bc=len(bcs_meta))
    for i, b in enumerate(bcs_meta):
        for k, v in b.items():
            out[f"bc{i}_{k}"] = v
    out["f_init"] = f_init
    out["f_final"] = np.asarray(f_0)
    out["bc_mask"] = np.asarray(bc_mask)
    out["missing_bits"] = pack_bits(missing)
    out["rho"] = np.asarray(rho)
    out["u"] = np.asarray(u)
    if force_bc is not None:  # reference MomentumTransfer (operator/force/momentum_transfer.py:51-90) on the final state
        from xlb.operator.force.momentum_transfer import MomentumTransfer

        out["force"] = np.asarray(MomentumTransfer(force_bc)(f_0, f_1, bc_mask, missing))
        out["force_bc"] = np.int64(stepper.boundary_conditions.index(force_bc))
    assert not np.isnan(out["f_final"].astype(np.float64)).any(), name
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {out['f_final'].shape} {out['f_final'].dtype} -> {os.path.getsize(path) / 1024:.0f} KiB")


def cavity(name, lattice, policy, n, steps, collision="BGK", omega=1.0):
    """examples/performance/mlups_3d.py:45-63 (3-D) / examples/cfd/lid_driven_cavity_2d.py (2-D) set-up."""
    vs, pp = init(lattice, policy)
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
    run_and_save(name, dict(lattice=lattice, policy=policy, collision=collision, shape=np.array(shape)), stepper, meta, steps, omega)


def sphere(name, lattice, policy, shape, steps, collision, omega=1.6, outlet="outflow", inlet="regularized"):
    """examples/cfd/flow_past_sphere_3d.py:41-109 geometry (walls Fullway, Poiseuille velocity inlet, outlet, Halfway sphere)."""
    vs, pp = init(lattice, policy)
    u_max = 0.04
    grid = grid_factory(shape)
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    inlet_idx, outlet_idx = bne["left"], bne["right"]
    walls = [box["bottom"][i] + box["top"][i] + box["front"][i] + box["back"][i] for i in range(3)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    r = max(2, shape[1] // 6)
    X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    ind = np.where((X - shape[0] // 6) ** 2 + (Y - shape[1] // 2) ** 2 + (Z - shape[2] // 2) ** 2 < r**2)
    sph = [tuple(ind[i]) for i in range(3)]
    H_y, H_z = float(shape[1] - 1), float(shape[2] - 1)

    def profile():
        y, z = jnp.arange(shape[1]), jnp.arange(shape[2])
        Yj, Zj = jnp.meshgrid(y, z, indexing="ij")
        yc, zc = Yj - (H_y / 2.0), Zj - (H_z / 2.0)
        r2 = (2.0 * yc / H_y) ** 2.0 + (2.0 * zc / H_z) ** 2.0
        ux = u_max * jnp.maximum(0.0, 1.0 - r2)
        return jnp.stack([ux, jnp.zeros_like(ux), jnp.zeros_like(ux)])

    bc_walls = FullwayBounceBackBC(indices=walls)
    In = RegularizedBC if inlet == "regularized" else ZouHeBC
    bc_in = In("velocity", profile=profile, indices=inlet_idx)
    if outlet == "outflow":
        bc_out = ExtrapolationOutflowBC(indices=outlet_idx)
        out_meta = dict(kind="outflow")
    elif outlet == "pressure":
        bc_out = ZouHeBC("pressure", prescribed_value=1.0, indices=outlet_idx)
        out_meta = dict(kind="zouhe", bc_type="pressure", prescribed=np.float64(1.0))
    elif outlet == "regularized_pressure":
        bc_out = RegularizedBC("pressure", prescribed_value=1.0, indices=outlet_idx)
        out_meta = dict(kind="regularized", bc_type="pressure", prescribed=np.float64(1.0))
    else:
        bc_out = DoNothingBC(indices=outlet_idx)
        out_meta = dict(kind="donothing")
    bc_sph = HalfwayBounceBackBC(indices=sph)
    bcs = [bc_walls, bc_in, bc_out, bc_sph]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=collision)
    # the prescribed inlet profile as the reference evaluates it (stored by aux_data_init on the JAX path)
    pv = np.asarray(profile())
    out_meta.update(id=bc_out.id, indices=np.array(outlet_idx))
    meta = [
        dict(kind="fullway", id=bc_walls.id, indices=np.array(walls)),
        dict(kind=inlet, id=bc_in.id, indices=np.array(inlet_idx), bc_type="velocity", prescribed=pv),
        out_meta,
        dict(kind="halfway", id=bc_sph.id, indices=np.array(sph)),
    ]
    run_and_save(name, dict(lattice=lattice, policy=policy, collision=collision, shape=np.array(shape)), stepper, meta, steps, omega, force_bc=bc_sph)


def channel2d(name, policy, shape, steps, collision, omega, outlet="outflow", inlet="regularized"):
    """2-D channel past a cylinder (D2Q9): Fullway walls, parabolic velocity inlet, outlet, Halfway cylinder — the 2-D
    counterpart of the sphere case (same BC classes as examples/cfd/flow_past_sphere_3d.py)."""
    vs, pp = init("D2Q9", policy)
    grid = grid_factory(shape)
    box, bne = grid.bounding_box_indices(), grid.bounding_box_indices(remove_edges=True)
    inlet_idx, outlet_idx = bne["left"], bne["right"]
    walls = [box["bottom"][i] + box["top"][i] for i in range(2)]
    walls = np.unique(np.array(walls), axis=-1).tolist()
    X, Y = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    ind = np.where((X - shape[0] // 5) ** 2 + (Y - shape[1] // 2) ** 2 < (shape[1] // 6) ** 2)
    cyl = [tuple(ind[i]) for i in range(2)]
    H = float(shape[1] - 1)

    def profile():
        y = jnp.arange(shape[1])
        ux = 0.04 * jnp.maximum(0.0, 1.0 - (2.0 * (y - H / 2.0) / H) ** 2.0)
        return jnp.stack([ux, jnp.zeros_like(ux)])

    bc_walls = FullwayBounceBackBC(indices=walls)
    In = RegularizedBC if inlet == "regularized" else ZouHeBC
    bc_in = In("velocity", profile=profile, indices=inlet_idx)
    if outlet == "outflow":
        bc_out, out_meta = ExtrapolationOutflowBC(indices=outlet_idx), dict(kind="outflow")
    else:
        bc_out = ZouHeBC("pressure", prescribed_value=1.0, indices=outlet_idx)
        out_meta = dict(kind="zouhe", bc_type="pressure", prescribed=np.float64(1.0))
    bc_cyl = HalfwayBounceBackBC(indices=cyl)
    bcs = [bc_walls, bc_in, bc_out, bc_cyl]
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=bcs, collision_type=collision)
    out_meta.update(id=bc_out.id, indices=np.array(outlet_idx))
    meta = [
        dict(kind="fullway", id=bc_walls.id, indices=np.array(walls)),
        dict(kind=inlet, id=bc_in.id, indices=np.array(inlet_idx), bc_type="velocity", prescribed=np.asarray(profile())),
        out_meta,
        dict(kind="halfway", id=bc_cyl.id, indices=np.array(cyl)),
    ]
    run_and_save(name, dict(lattice="D2Q9", policy=policy, collision=collision, shape=np.array(shape)), stepper, meta, steps, omega, force_bc=bc_cyl)


def periodic(name, lattice, policy, shape, steps, collision, omega, force=None):
    """Fully periodic box from a seeded random velocity field (as examples/cfd/turbulent_channel_3d.py:130-135 seeds u)."""
    vs, pp = init(lattice, policy)
    grid = grid_factory(shape)
    rng = np.random.default_rng(0)
    u0 = (1e-2 * rng.standard_normal((vs.d,) + tuple(shape))).astype(pp.compute_precision.jax_dtype)
    rho0 = (1.0 + 1e-3 * rng.standard_normal((1,) + tuple(shape))).astype(pp.compute_precision.jax_dtype)

    def initializer(grid, velocity_set, precision_policy, compute_backend):
        f = initialize_eq(None, grid, velocity_set, precision_policy, compute_backend, rho=jnp.array(rho0), u=jnp.array(u0))
        return precision_policy.cast_to_store_jax(f)

    kw = {} if force is None else dict(force_vector=jnp.array(force, dtype=pp.compute_precision.jax_dtype))
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=[], collision_type=collision, **kw)
    meta = dict(lattice=lattice, policy=policy, collision=collision, shape=np.array(shape), rho0=rho0, u0=u0)
    if force is not None:
        meta["force_vector"] = np.asarray(force, dtype=np.float64)
    run_and_save(name, meta, stepper, [], steps, omega, initializer)


def operator_vectors():
    """Single-operator outputs of the reference on seeded inputs (Stream, Equilibrium, Macroscopic, SecondMoment, BGK, KBC)."""
    from xlb.operator.stream import Stream
    from xlb.operator.equilibrium import QuadraticEquilibrium
    from xlb.operator.collision import BGK, KBC
    from xlb.operator.macroscopic import SecondMoment

    out = {}
    for lattice, shape in (("D2Q9", (12, 10)), ("D3Q19", (8, 7, 6)), ("D3Q27", (8, 7, 6))):
        vs, pp = init(lattice, "FP32FP32")
        rng = np.random.default_rng(1)
        rho = (1.0 + 0.05 * rng.standard_normal((1,) + shape)).astype(np.float32)
        u = (0.05 * rng.standard_normal((vs.d,) + shape)).astype(np.float32)
        feq = QuadraticEquilibrium()(jnp.array(rho), jnp.array(u))
        f = np.asarray(feq) * (1.0 + 0.02 * rng.standard_normal((vs.q,) + shape)).astype(np.float32)
        f = f.astype(np.float32)
        r2, u2 = Macroscopic()(jnp.array(f))
        feq2 = QuadraticEquilibrium()(r2, u2)
        out.update({f"{lattice}_rho": rho, f"{lattice}_u": u, f"{lattice}_feq": np.asarray(feq), f"{lattice}_f": f})
        out.update({f"{lattice}_stream": np.asarray(Stream()(jnp.array(f))), f"{lattice}_rho2": np.asarray(r2), f"{lattice}_u2": np.asarray(u2)})
        out[f"{lattice}_pi"] = np.asarray(SecondMoment()(jnp.array(f)))
        out[f"{lattice}_bgk"] = np.asarray(BGK()(jnp.array(f), feq2, r2, u2, 1.3))
        if lattice != "D3Q19":
            out[f"{lattice}_kbc"] = np.asarray(KBC()(jnp.array(f), jnp.array(feq2), jnp.array(r2), jnp.array(u2), 1.7))
    path = os.path.join(HERE, "operators.npz")
    np.savez_compressed(path, **out)
    print(f"operators: {len(out)} arrays -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    operator_vectors()
    cavity("cavity_d3q19_bgk_fp32", "D3Q19", "FP32FP32", 16, 60)
    cavity("cavity_d3q19_bgk_fp32fp16", "D3Q19", "FP32FP16", 16, 60)
    cavity("cavity_d3q19_bgk_fp64fp32", "D3Q19", "FP64FP32", 16, 60)
    cavity("cavity_d3q27_kbc_fp32", "D3Q27", "FP32FP32", 14, 40, collision="KBC", omega=1.7)
    cavity("cavity_d2q9_bgk_fp32", "D2Q9", "FP32FP32", 32, 100, omega=1.5)
    cavity("cavity_d2q9_kbc_fp32", "D2Q9", "FP32FP32", 32, 100, collision="KBC", omega=1.8)
    sphere("sphere_d3q27_kbc_fp32", "D3Q27", "FP32FP32", (40, 16, 16), 40, "KBC")
    sphere("sphere_d3q19_bgk_fp32", "D3Q19", "FP32FP32", (40, 16, 16), 40, "BGK")
    sphere("sphere_d3q19_bgk_zouhe_pressure_fp32", "D3Q19", "FP32FP32", (32, 14, 14), 30, "BGK", omega=1.4, outlet="pressure", inlet="zouhe")
    sphere("sphere_d3q27_bgk_regpressure_fp64", "D3Q27", "FP64FP64", (32, 14, 14), 30, "BGK", omega=1.4, outlet="regularized_pressure")
    sphere("sphere_d3q19_bgk_donothing_fp32", "D3Q19", "FP32FP32", (32, 14, 14), 30, "BGK", omega=1.2, outlet="donothing")
    periodic("periodic_d3q19_bgk_fp32", "D3Q19", "FP32FP32", (12, 10, 8), 30, "BGK", 1.5)
    periodic("periodic_d3q27_kbc_fp32", "D3Q27", "FP32FP32", (12, 10, 8), 30, "KBC", 1.8)
    sphere("sphere_d3q19_bgk_fp32fp16", "D3Q19", "FP32FP16", (32, 14, 14), 30, "BGK", omega=1.5)
    sphere("sphere_d3q27_kbc_fp64fp32", "D3Q27", "FP64FP32", (32, 14, 14), 30, "KBC", omega=1.6)
    sphere("sphere_d3q27_kbc_zouhe_pressure_fp64", "D3Q27", "FP64FP64", (28, 12, 12), 25, "KBC", omega=1.7, outlet="pressure", inlet="zouhe")
    channel2d("channel2d_d2q9_bgk_outflow_fp32", "FP32FP32", (60, 24), 60, "BGK", 1.6)
    channel2d("channel2d_d2q9_bgk_zouhe_pressure_fp32", "FP32FP32", (60, 24), 60, "BGK", 1.5, outlet="pressure", inlet="zouhe")
    # body force (ForcedCollision + ExactDifference, as examples/cfd/turbulent_channel_3d.py drives its channel): target for round 2
    periodic("periodic_d3q19_bgk_forced_fp32", "D3Q19", "FP32FP32", (12, 10, 8), 30, "BGK", 1.5, force=(1e-5, 0.0, 0.0))
