"""BASELINE.json's full sizes (512^3): size-independent properties instead of an oracle run.
 * closed periodic box: total mass is conserved by stream + BGK collide (to fp32 rounding),
 * every vector width of the kernel gives the same field,
 * a uniform rest state is a fixed point of the step (idempotence), also with fp16 storage."""

import numpy as np
import pytest
import torch

import xlb_b200 as xlb
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.grid import grid_factory
from xlb_b200.helper import initialize_eq
from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

pytestmark = pytest.mark.gpu
N = 512


def make(policy, lattice="D3Q19", v=0, taylor_green=True):
    pp = xlb.PrecisionPolicy[policy]
    be = ComputeBackend.WARP
    vs = getattr(xlb.velocity_set, lattice)(precision_policy=pp, compute_backend=be)
    xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
    grid = grid_factory((N, N, N))
    stepper = IncompressibleNavierStokesStepper(grid=grid, boundary_conditions=[], collision_type="BGK", cells_per_thread=v)

    def initializer(grid, velocity_set, precision_policy, compute_backend):
        f = grid.create_field(cardinality=velocity_set.q)
        rho = grid.create_field(cardinality=1, fill_value=1.0, dtype=precision_policy.compute_precision)
        u = grid.create_field(cardinality=3, fill_value=0.0, dtype=precision_policy.compute_precision)
        if taylor_green:
            x = torch.arange(N, device=grid.device, dtype=torch.float32) * (2 * np.pi / N)
            X, Y, Z = torch.meshgrid(x, x, x, indexing="ij")
            u[0] = 0.02 * torch.sin(X) * torch.cos(Y) * torch.cos(Z)
            u[1] = -0.02 * torch.cos(X) * torch.sin(Y) * torch.cos(Z)
            del X, Y, Z
        return initialize_eq(f, grid, velocity_set, precision_policy, compute_backend, rho=rho, u=u)

    return stepper, stepper.prepare_fields(initializer=initializer)


def run(stepper, fields, steps, omega=1.7):
    f_0, f_1, bc_mask, missing_mask = fields
    for i in range(steps):
        f_0, f_1 = stepper(f_0, f_1, bc_mask, missing_mask, omega, i)
        f_0, f_1 = f_1, f_0
    return f_0


def total_mass(f):
    return float(sum(f[l].sum(dtype=torch.float64) for l in range(f.shape[0])))


def test_mass_conservation_512():
    stepper, fields = make("FP32FP32")
    m0 = total_mass(fields[0])
    f = run(stepper, fields, 20)
    assert abs(total_mass(f) - m0) / m0 < 1e-6
    assert bool(torch.isfinite(f).all())


def test_vector_widths_agree_512():
    ref = None
    for v in (1, 2, 4):
        stepper, fields = make("FP32FP32", v=v)
        f = run(stepper, fields, 5)
        if ref is None:
            ref = f.clone()
        else:
            assert float((f - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        del stepper, fields, f
        torch.cuda.empty_cache()


@pytest.mark.parametrize("policy", ["FP32FP32", "FP32FP16"])
def test_rest_state_is_a_fixed_point_512(policy):
    stepper, fields = make(policy, taylor_green=False)
    before = fields[0].clone()
    f = run(stepper, fields, 4, omega=1.0)
    tol = 1e-6 if policy == "FP32FP32" else 1e-3
    assert float((f.float() - before.float()).abs().max()) <= tol * float(before.float().abs().max())
