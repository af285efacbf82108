"""Duplicate-index detection across boundary conditions (reference: xlb/helper/check_boundary_overlaps.py:5-24):
an error on the WARP convention (the masker would be order dependent), a warning on the JAX convention."""

import numpy as np

from xlb_b200.compute_backend import ComputeBackend


def _has_duplicates(index_rows) -> bool:
    arr = np.asarray(index_rows)
    if arr.size == 0:
        return False
    return np.unique(arr, axis=-1).shape[-1] != arr.shape[-1]


def check_bc_overlaps(bclist, dim, compute_backend):
    merged = [[] for _ in range(dim)]
    for bc in bclist:
        if bc.indices is None:
            continue
        if _has_duplicates(bc.indices):
            if compute_backend == ComputeBackend.WARP:
                raise ValueError(f"Boundary condition {bc.__class__.__name__} has duplicate indices!")
            print(f"WARNING: there are duplicate indices in {bc.__class__.__name__} and hence the order in bc list matters!")
        for d in range(dim):
            merged[d] += list(bc.indices[d])
    if _has_duplicates(merged):
        if compute_backend == ComputeBackend.WARP:
            raise ValueError("Boundary condition list containes duplicate indices!")
        print("WARNING: there are duplicate indices in the boundary condition list and hence the order in this list matters!")
