"""D3Q19 lattice (reference ordering: xlb/velocity_set/d3q19.py:19-27)."""

import itertools

import numpy as np

from xlb_b200.velocity_set.velocity_set import VelocitySet, _weights_by_speed


class D3Q19(VelocitySet):
    lattice_code = 1

    def __init__(self, precision_policy=None, compute_backend=None):
        c = np.array([v for v in itertools.product((0, -1, 1), repeat=3) if sum(map(abs, v)) <= 2]).T
        w = _weights_by_speed(c, {0: 1.0 / 3.0, 1: 1.0 / 18.0, 2: 1.0 / 36.0})
        super().__init__(3, 19, c, w, precision_policy=precision_policy, compute_backend=compute_backend)
