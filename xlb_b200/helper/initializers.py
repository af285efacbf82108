"""Equilibrium initialisation f = feq(rho, u); default rho = 1, u = 0 (reference: xlb/helper/initializers.py:5-20)."""

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.equilibrium import QuadraticEquilibrium


def initialize_eq(f, grid, velocity_set, precision_policy, compute_backend, rho=None, u=None):
    if rho is None:
        rho = grid.create_field(cardinality=1, fill_value=1.0, dtype=precision_policy.compute_precision)
    if u is None:
        u = grid.create_field(cardinality=velocity_set.d, fill_value=0.0, dtype=precision_policy.compute_precision)
    equilibrium = QuadraticEquilibrium(velocity_set, precision_policy, compute_backend)
    if compute_backend == ComputeBackend.JAX:
        f = equilibrium(rho, u).to(precision_policy.store_precision.torch_dtype)
    else:
        f = equilibrium(rho, u, f)
    del rho, u
    return f
