"""D3Q27 lattice (reference ordering: xlb/velocity_set/d3q27.py:19-29)."""

import itertools

import numpy as np

from xlb_b200.velocity_set.velocity_set import VelocitySet, _weights_by_speed


class D3Q27(VelocitySet):
    lattice_code = 2

    def __init__(self, precision_policy=None, compute_backend=None):
        c = np.array(list(itertools.product((0, -1, 1), repeat=3))).T
        w = _weights_by_speed(c, {0: 8.0 / 27.0, 1: 2.0 / 27.0, 2: 1.0 / 54.0, 3: 1.0 / 216.0})
        super().__init__(3, 27, c, w, precision_policy=precision_policy, compute_backend=compute_backend)
