"""Host-side counterparts of the reference's Warp BC helpers (helper_functions_bc.py:36-122).

In the reference these are ``@wp.func`` device functions compiled into every BC kernel.  Here the device versions are
C++ (bc_normal / bc_fsum / bc_bounceback_nonequilibrium / bc_regularize in xlb_b200/csrc/lbm_math.cuh); this class keeps
the name importable and offers the one helper the host needs: outward normals of boundary cells from the missing mask,
used to turn a prescribed velocity VECTOR (JAX convention) into the per-cell normal magnitude the kernels consume.
"""

import numpy as np


class HelperFunctionsBC(object):
    def __init__(self, velocity_set=None, precision_policy=None, compute_backend=None):
        from xlb_b200.default_config import DefaultConfig

        self.velocity_set = velocity_set or DefaultConfig.velocity_set
        self.precision_policy = precision_policy or DefaultConfig.default_precision_policy
        self.compute_backend = compute_backend or DefaultConfig.default_backend

    def get_normal_vectors(self, missing_cells: np.ndarray) -> np.ndarray:
        """missing_cells: bool (q, n).  Returns int (d, n): minus the FIRST missing axis-aligned direction in index
        order (reference: helper_functions_bc.py:75-86); zero where no main direction is missing."""
        vs = self.velocity_set
        n = np.zeros((vs.d, missing_cells.shape[1]), dtype=np.int64)
        found = np.zeros(missing_cells.shape[1], dtype=bool)
        for l in vs.main_indices:
            sel = missing_cells[l] & ~found
            n[:, sel] = -vs._c[:, l : l + 1]
            found |= sel
        return n
