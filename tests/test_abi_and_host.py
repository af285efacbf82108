"""CPU-only checks: the C-ABI library loads and exports every symbol include/xlb_b200.h declares; the lattice tables
compiled into the kernels equal the host tables and the oracle's; host-side mirrors of the reference interface
(grid, registry, overlap check, precision policy); operators refuse CPU tensors loudly (no fallback)."""

import os
import re

import numpy as np
import pytest
import torch

import xlb_b200 as xlb
from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.grid import grid_factory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def init(lattice="D3Q19", backend=ComputeBackend.WARP, policy=xlb.PrecisionPolicy.FP32FP32):
    vs = getattr(xlb.velocity_set, lattice)(precision_policy=policy, compute_backend=backend)
    xlb.init(velocity_set=vs, default_backend=backend, default_precision_policy=policy)
    return vs


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "xlb_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(xlbn_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(native.SIGNATURES), "ctypes signatures and header disagree"
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by libxlb_b200.so"
    assert lib.xlbn_version() == 100


@pytest.mark.parametrize("lattice", ["D2Q9", "D3Q19", "D3Q27"])
def test_kernel_lattice_tables_match_host_and_oracle(lattice):
    from oracle import lbm_numpy as O

    vs = init(lattice)
    c, w, opp = native.lattice_tables(vs.lattice_code)
    lat = O.Lattice(lattice)
    assert np.array_equal(c[: vs.d], vs._c) and np.array_equal(c[: vs.d], lat.c)
    assert np.all(c[vs.d :] == 0)
    assert np.allclose(w, vs._w, rtol=0, atol=0) and np.allclose(w, lat.w, rtol=0, atol=0)
    assert np.array_equal(opp, vs.opp_indices) and np.array_equal(opp, lat.opp)
    assert np.array_equal(vs._cc, lat.cc) and np.allclose(vs._qi, lat.qi, rtol=0, atol=0)
    assert np.array_equal(vs.main_indices, lat.main) and np.array_equal(vs.right_indices, lat.right) and np.array_equal(vs.left_indices, lat.left)


def test_velocity_set_orders_are_the_reference_orders():
    vs = init("D3Q19")
    assert vs.opp_indices.tolist() == [0, 2, 1, 6, 8, 7, 3, 5, 4, 14, 16, 15, 18, 17, 9, 11, 10, 13, 12]
    assert vs.main_indices.tolist() == [1, 2, 3, 6, 9, 14] and vs.right_indices.tolist() == [14, 15, 16, 17, 18]
    vs = init("D3Q27")
    assert vs.opp_indices.tolist() == [0, 2, 1, 6, 8, 7, 3, 5, 4, 18, 20, 19, 24, 26, 25, 21, 23, 22, 9, 11, 10, 15, 17, 16, 12, 14, 13]
    vs = init("D2Q9")
    assert vs.opp_indices.tolist() == [0, 2, 1, 6, 5, 4, 3, 8, 7] and vs.left_indices.tolist() == [4, 6, 8]
    assert np.allclose(vs._w, [4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 9, 1 / 36, 1 / 36])


@pytest.mark.parametrize("shape", [(7, 5), (6, 5, 4)])
@pytest.mark.parametrize("remove_edges", [False, True])
def test_bounding_box_indices_match_reference_layout(shape, remove_edges):
    from oracle import lbm_numpy as O

    init("D2Q9" if len(shape) == 2 else "D3Q19")
    box = grid_factory(shape).bounding_box_indices(remove_edges=remove_edges)
    ref = O.bounding_box_indices(shape, remove_edges=remove_edges)
    assert set(box) == set(ref)
    for k in ref:
        assert isinstance(box[k], list) and np.array_equal(np.array(box[k]), ref[k]), k


def test_grid_fields_and_precision_policy():
    init("D3Q19", policy=xlb.PrecisionPolicy.FP32FP16)
    g = grid_factory((4, 5, 6))
    f = g.create_field(cardinality=19)
    assert tuple(f.shape) == (19, 4, 5, 6) and f.dtype == torch.float16 and f.is_contiguous()
    m = g.create_field(cardinality=1, dtype=xlb.Precision.UINT8)
    assert m.dtype == xlb.Precision.UINT8.wp_dtype == torch.uint8
    one = g.create_field(cardinality=1, dtype=xlb.Precision.FP64, fill_value=1.5)
    assert float(one.numpy().max()) == 1.5 and one.dtype == torch.float64
    init("D2Q9")
    assert tuple(grid_factory((4, 5), compute_backend=ComputeBackend.WARP).create_field(9).shape) == (9, 4, 5, 1)
    assert tuple(grid_factory((4, 5), compute_backend=ComputeBackend.JAX).create_field(9).shape) == (9, 4, 5)
    pp = xlb.PrecisionPolicy.FP64FP32
    assert pp.compute_precision == xlb.Precision.FP64 and pp.store_precision == xlb.Precision.FP32


def test_bc_registry_ids_follow_construction_order():
    from xlb_b200.operator.boundary_condition import EquilibriumBC, FullwayBounceBackBC, HalfwayBounceBackBC, RegularizedBC, ZouHeBC

    init("D3Q19")
    a = FullwayBounceBackBC(indices=[[0], [0], [0]])
    b = EquilibriumBC(rho=1.0, u=(0.0, 0.0, 0.0), indices=[[1], [1], [1]])
    c = RegularizedBC("velocity", prescribed_value=(0.1, 0.0, 0.0), indices=[[2], [2], [2]])
    assert (a.id, b.id, c.id) == (1, 2, 3)
    assert (a.needs_padding, b.needs_padding, c.needs_padding) == (False, False, True)
    assert c.needs_aux_init and c.needs_aux_recovery and c.num_of_aux_data == 1
    assert HalfwayBounceBackBC(indices=[[3], [3], [3]]).needs_padding
    with pytest.raises(ValueError):
        ZouHeBC("velocity", prescribed_value=(0.1, 0.1, 0.0))
    with pytest.raises(AssertionError):
        ZouHeBC("temperature")


def test_check_bc_overlaps():
    from xlb_b200.helper import check_bc_overlaps
    from xlb_b200.operator.boundary_condition import FullwayBounceBackBC

    init("D3Q19")
    a = FullwayBounceBackBC(indices=[[0, 1], [0, 1], [0, 1]])
    b = FullwayBounceBackBC(indices=[[1, 2], [1, 2], [1, 2]])
    with pytest.raises(ValueError):
        check_bc_overlaps([a, b], 3, ComputeBackend.WARP)
    check_bc_overlaps([a, b], 3, ComputeBackend.JAX)  # warns only
    check_bc_overlaps([a], 3, ComputeBackend.WARP)


def test_kbc_rejects_d3q19_like_the_reference():
    from xlb_b200.operator.collision import KBC

    init("D3Q19")
    with pytest.raises(NotImplementedError):
        KBC()


def test_operators_refuse_cpu_tensors_loudly():
    """No CPU / PyTorch fallback: a tensor that is not on a CUDA device is an error, never a silent slow path."""
    from xlb_b200.operator.equilibrium import QuadraticEquilibrium
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    vs = init("D3Q19")
    rho, u, f = torch.ones(1, 4, 4, 4), torch.zeros(3, 4, 4, 4), torch.zeros(19, 4, 4, 4)
    with pytest.raises(Exception, match="no CPU fallback"):
        QuadraticEquilibrium()(rho, u, f)
    g = grid_factory((4, 4, 4), device="cpu")
    stepper = IncompressibleNavierStokesStepper(grid=g, boundary_conditions=[])
    with pytest.raises(Exception, match="no CPU fallback"):
        stepper(f, f.clone(), torch.zeros(1, 4, 4, 4, dtype=torch.uint8), torch.zeros(19, 4, 4, 4, dtype=torch.bool), 1.0, 0)


def test_xlb_alias_package_is_the_same_module_tree():
    import xlb as alias
    from xlb.operator.stepper import IncompressibleNavierStokesStepper as A
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper as B

    assert alias is xlb and A is B
    import xlb.operator.boundary_condition.boundary_condition_registry as r1
    import xlb_b200.operator.boundary_condition.boundary_condition_registry as r2

    assert r1.boundary_condition_registry is r2.boundary_condition_registry


def test_zouhe_prescribed_values_both_conventions():
    """JAX convention (velocity vector / (d,ny,nz) profile) and Warp convention (normal magnitude / per-index callable)
    give the same per-cell scalar = -(u . n)."""
    from xlb_b200.operator.boundary_condition import RegularizedBC

    shape = (6, 5, 4)
    yy, zz = np.meshgrid(np.arange(1, 4), np.arange(1, 3), indexing="ij")
    cells = np.stack([np.zeros(yy.size, dtype=np.int64), yy.ravel(), zz.ravel()])  # x = 0 face
    vs = init("D3Q19", ComputeBackend.WARP)
    missing = np.zeros((19, cells.shape[1]), dtype=bool)
    missing[vs.right_indices] = True  # pulls from x-1 are missing on the x = 0 face
    prof = np.zeros((3, 5, 4))
    prof[0] = 0.01 * np.arange(20).reshape(5, 4)
    bc_w = RegularizedBC("velocity", profile=lambda index: [0.01 * (index[1] * 4 + index[2])])
    v_w = bc_w._prescribed_values_at(cells, missing, shape)
    init("D3Q19", ComputeBackend.JAX)
    bc_j = RegularizedBC("velocity", profile=lambda: prof)
    v_j = bc_j._prescribed_values_at(cells, missing, shape)
    assert np.allclose(v_w, v_j) and np.allclose(v_w, prof[0][cells[1], cells[2]])
    bc_c = RegularizedBC("velocity", prescribed_value=(0.03, 0.0, 0.0))
    assert np.allclose(bc_c._prescribed_values_at(cells, missing, shape), 0.03)
    init("D3Q19", ComputeBackend.WARP)
    bc_c = RegularizedBC("velocity", prescribed_value=(0.03, 0.0, 0.0))
    assert np.allclose(bc_c._prescribed_values_at(cells, missing, shape), 0.03)


def test_new_entry_points_validate_their_arguments_without_a_device():
    """Argument-error paths of the entry points added for SURVEY §8f N3 / N4: they return before touching CUDA, so the ctypes
    signatures (argument count and order) are exercised here on the CPU."""
    import ctypes as C

    from xlb_b200 import native

    lib = native.lib()
    err = lambda: lib.xlbn_last_error().decode()
    dims = native.int3((4, 4, 4))
    one = C.c_void_p(1)  # non-NULL, never dereferenced on these paths
    # xlbn_collide_ext: NULL arrays, unknown / unsupported collisions, forced without u / force
    assert lib.xlbn_collide_ext(native.D3Q19, native.BGK, native.F32, None, 1, None, 1, None, 1, None, 0, None, 0, 1.0, None, 0.17, dims, None) < 0 and "NULL array" in err()
    assert lib.xlbn_collide_ext(native.D3Q19, 3, native.F32, one, 1, one, 1, one, 1, None, 0, None, 0, 1.0, None, 0.17, dims, None) < 0 and "unknown collision" in err()
    assert lib.xlbn_collide_ext(native.D3Q19, native.KBC, native.F32, one, 1, one, 1, one, 1, one, 1, None, 0, 1.0, None, 0.17, dims, None) < 0 and "D3Q19" in err()
    dims2 = native.int3((4, 4, 1))
    assert lib.xlbn_collide_ext(native.D2Q9, native.SMAGORINSKY_LES_BGK, native.F32, one, 1, one, 1, one, 1, None, 0, None, 0, 1.0, None, 0.17, dims2, None) < 0 and "3-D velocity sets only" in err()
    assert lib.xlbn_collide_ext(native.D3Q27, native.BGK | native.COLLISION_FORCED, native.F32, one, 1, one, 1, one, 1, one, 1, None, 0, 1.0, None, 0.17, dims, None) < 0 and "needs u and the force" in err()
    # xlbn_exact_difference: NULL argument
    assert lib.xlbn_exact_difference(native.D3Q19, native.F32, one, 1, one, 1, one, 1, one, 1, one, 1, None, dims, None) < 0 and "NULL argument" in err()
    # xlbn_mask_mesh: 2-D lattice, NULL, bad id, bad edge test
    assert lib.xlbn_mask_mesh(native.D2Q9, one, 1, 1, 0, dims2, one, one, one, None) < 0 and "3-D lattices only" in err()
    assert lib.xlbn_mask_mesh(native.D3Q19, None, 1, 1, 0, dims, one, one, one, None) < 0 and "NULL argument" in err()
    assert lib.xlbn_mask_mesh(native.D3Q19, one, 1, 255, 0, dims, one, one, one, None) < 0 and "id 255" in err()
    assert lib.xlbn_mask_mesh(native.D3Q19, one, 1, 1, 7, dims, one, one, one, None) < 0 and "edge_test 7" in err()
    # stepper setters / prepare / halo liveness calls on a NULL handle
    assert lib.xlbn_stepper_prepare(None, 1.0, None) < 0 and "NULL stepper" in err()
    assert lib.xlbn_halo_set_timeout(None, 1.0) < 0 and lib.xlbn_halo_timed_out(None) < 0
    assert lib.xlbn_stepper_set_force(None, (C.c_double * 3)(1e-5, 0.0, 0.0)) < 0 and "NULL stepper" in err()
    assert lib.xlbn_stepper_set_smagorinsky(None, 0.17) < 0 and "NULL stepper" in err()
    # stepper_create argument checks for the new options (they fail before cudaMalloc)
    out = C.c_void_p()
    desc = native.StepperDesc(lattice=native.D2Q9, collision=native.SMAGORINSKY_LES_BGK, compute_dtype=native.F32, store_dtype=native.F32, n_bc=0, cells_per_thread=0, bcs=None)
    assert lib.xlbn_stepper_create(C.byref(desc), C.byref(out)) < 0 and "3-D velocity sets only" in err()
    desc = native.StepperDesc(lattice=native.D3Q19, collision=native.BGK, compute_dtype=native.F32, store_dtype=native.F32, n_bc=0, cells_per_thread=301, bcs=None)
    assert lib.xlbn_stepper_create(C.byref(desc), C.byref(out)) < 0 and "selects a KBC formulation" in err()
    desc = native.StepperDesc(lattice=native.D3Q19, collision=5, compute_dtype=native.F32, store_dtype=native.F32, n_bc=0, cells_per_thread=0, bcs=None)
    assert lib.xlbn_stepper_create(C.byref(desc), C.byref(out)) < 0 and "unknown collision" in err()
