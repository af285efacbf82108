"""Duplicate-index detection across boundary conditions (reference: xlb/helper/check_boundary_overlaps.py:5-24):
an error on the WARP convention (the masker would be order dependent), a warning on the JAX convention."""

import numpy as np

from xlb_b200.compute_backend import ComputeBackend


def _has_duplicates(index_rows) -> bool:
    arr = np.asarray(index_rows)
    return arr.size > 0 and np.unique(arr, axis=-1).shape[-1] != arr.shape[-1]


def _report(compute_backend, error, warning):
    if compute_backend == ComputeBackend.WARP:
        raise ValueError(error)
    print("WARNING: " + warning)


def check_bc_overlaps(bclist, dim, compute_backend):
    with_indices = [bc for bc in bclist if bc.indices is not None]
    for bc in with_indices:
        name = bc.__class__.__name__
        if _has_duplicates(bc.indices):
            _report(compute_backend, f"Boundary condition {name} has duplicate indices!", f"there are duplicate indices in {name} and hence the order in bc list matters!")
    merged = [[i for bc in with_indices for i in bc.indices[d]] for d in range(dim)]
    if _has_duplicates(merged):
        _report(
            compute_backend,
            "Boundary condition list containes duplicate indices!",
            "there are duplicate indices in the boundary condition list and hence the order in this list matters!",
        )
