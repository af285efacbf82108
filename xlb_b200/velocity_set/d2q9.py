"""D2Q9 lattice (reference ordering is hand-listed: xlb/velocity_set/d2q9.py:18-21)."""

import numpy as np

from xlb_b200.velocity_set.velocity_set import VelocitySet, _weights_by_speed


class D2Q9(VelocitySet):
    lattice_code = 0

    def __init__(self, precision_policy=None, compute_backend=None):
        c = np.array([(0, 0), (0, 1), (0, -1), (1, 0), (-1, 1), (1, -1), (-1, 0), (1, 1), (-1, -1)]).T
        w = _weights_by_speed(c, {0: 4.0 / 9.0, 1: 1.0 / 9.0, 2: 1.0 / 36.0})
        super().__init__(2, 9, c, w, precision_policy=precision_policy, compute_backend=compute_backend)
