"""Lattice velocity sets.

The index ORDER of the discrete velocities is part of the bit-exact contract
with the reference: D3Q19/D3Q27 enumerate
``itertools.product([0, -1, 1], repeat=3)`` (reference: xlb/velocity_set/d3q19.py:19,
d3q27.py:19) and D2Q9 is hand-listed (d2q9.py:18-21).  The derived tables
(opposites, second-moment products ``cc``, regularisation tensor ``qi``, the
axis-aligned ``main_indices`` and the x-face ``left/right_indices`` used by the
halo exchange) follow xlb/velocity_set/velocity_set.py:55-221.

The same tables are compiled into the CUDA library as constexpr arrays
(xlb_b200/csrc/lattice.cuh); tests/test_native_abi.py checks both agree.
"""

import math

import numpy as np


class VelocitySet(object):
    """Base class: d, q, c (d,q) int, w (q,) and derived index tables."""

    # lattice code of the C ABI (include/xlb_b200.h, xlbn_lattice)
    lattice_code = -1

    def __init__(self, d, q, c, w, precision_policy, compute_backend):
        self.d = d
        self.q = q
        self.precision_policy = precision_policy
        self.compute_backend = compute_backend

        c = np.asarray(c, dtype=np.int32)
        assert c.shape == (d, q)
        compute_np = precision_policy.compute_precision.np_dtype if precision_policy is not None else np.float64

        # numpy masters (float64), as the reference keeps under _c/_w/...
        self._c = c
        self._w = np.asarray(w, dtype=np.float64)
        self._opp_indices = self._opposites(c)
        self._cc = self._second_order_products(c)
        self._c_float = c.astype(np.float64)
        self._qi = self._regularisation_tensor(self._cc, d)

        # backend-facing copies in the compute dtype
        self.c = self._c
        self.w = self._w.astype(compute_np)
        self.opp_indices = self._opp_indices
        self.cc = self._cc.astype(compute_np)
        self.c_float = self._c_float.astype(compute_np)
        self.qi = self._qi.astype(compute_np)

        self.cs = compute_np(math.sqrt(3.0) / 3.0)
        self.cs2 = compute_np(1.0 / 3.0)
        self.inv_cs2 = compute_np(3.0)

        speed = np.abs(c).sum(axis=0)
        self.main_indices = np.nonzero(speed == 1)[0]
        self.right_indices = np.nonzero(c[0] == 1)[0]
        self.left_indices = np.nonzero(c[0] == -1)[0]

    @staticmethod
    def _opposites(c):
        lookup = {tuple(v): i for i, v in enumerate(c.T.tolist())}
        return np.array([lookup[tuple((-v).tolist())] for v in c.T], dtype=np.int32)

    @staticmethod
    def _second_order_products(c):
        d, q = c.shape
        pairs = [(a, b) for a in range(d) for b in range(a, d)]  # xx,xy,xz,yy,yz,zz / xx,xy,yy
        cc = np.zeros((q, len(pairs)))
        for t, (a, b) in enumerate(pairs):
            cc[:, t] = c[a] * c[b]
        return cc

    @staticmethod
    def _regularisation_tensor(cc, d):
        # Q_i = c_i c_i - cs^2 I, off-diagonals doubled (symmetric tensor stored as a vector)
        if d == 3:
            diagonal, off = (0, 3, 5), (1, 2, 4)
        elif d == 2:
            diagonal, off = (0, 2), (1,)
        else:
            raise ValueError(f"dim = {d} not supported")
        qi = cc.copy()
        qi[:, diagonal] -= 1.0 / 3.0
        qi[:, off] *= 2.0
        return qi

    def __repr__(self):
        return "D{}Q{}".format(self.d, self.q)

    __str__ = __repr__


def _weights_by_speed(c, table):
    return np.array([table[int(s)] for s in np.abs(c).sum(axis=0)], dtype=np.float64)
