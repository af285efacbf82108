"""Equilibrium initialisation f = feq(rho, u); default rho = 1, u = 0 (reference: xlb/helper/initializers.py:5-20)."""

from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.equilibrium import QuadraticEquilibrium


def initialize_eq(f, grid, velocity_set, precision_policy, compute_backend, rho=None, u=None):
    """JAX convention: returns a new field in the store dtype; WARP convention: fills and returns `f`."""
    compute = precision_policy.compute_precision
    moments = []
    for given, cardinality, fill in ((rho, 1, 1.0), (u, velocity_set.d, 0.0)):
        moments.append(given if given is not None else grid.create_field(cardinality=cardinality, fill_value=fill, dtype=compute))
    feq = QuadraticEquilibrium(velocity_set, precision_policy, compute_backend)
    if compute_backend == ComputeBackend.JAX:
        return feq(*moments).to(precision_policy.store_precision.torch_dtype)
    return feq(*moments, f)
