"""Zou-He boundary condition (velocity or pressure type): non-equilibrium bounce-back of the unknown populations.

Reference: xlb/operator/boundary_condition/bc_zouhe.py — ctor and prescribed-value handling L39-128, JAX L130-269,
Warp functionals L279-343.

Prescribed values.  The kernels consume ONE scalar per boundary cell, kept in the populations' rest slot of the
second buffer exactly like the reference's Warp path (``f_1[0, cell]``, boundary_condition.py:151): the magnitude of
the normal velocity (``u = -value * n``, bc_zouhe.py:302-303) or the density.  Both reference conventions are accepted:

* WARP: ``prescribed_value`` = vector with one non-zero entry (its value is taken) or a scalar density;
  ``profile(index)`` = callable returning a length-1 vector for one cell index (a ``@wp.func`` in reference scripts).
  It is evaluated on the host, vectorised over all BC cells when the callable allows it.
* JAX: ``prescribed_value`` = velocity vector / density; ``profile()`` = callable without arguments returning an array
  broadcastable over the boundary face ((d, ny, nz)-like, bc_zouhe.py:143-183).  The vector is projected on the cell's
  outward normal: value = -(u . n).
"""

from typing import Tuple, Union

import numpy as np

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.boundary_condition.boundary_condition import BoundaryCondition, ImplementationStep
from xlb_b200.operator.boundary_condition.helper_functions_bc import HelperFunctionsBC
from xlb_b200.operator.equilibrium import QuadraticEquilibrium
from xlb_b200.operator.operator import Operator


class ZouHeBC(BoundaryCondition):
    _kind_by_type = {"velocity": native.BC_ZOUHE_VELOCITY, "pressure": native.BC_ZOUHE_PRESSURE}

    def __init__(
        self,
        bc_type,
        profile=None,
        prescribed_value: Union[float, Tuple[float, ...], np.ndarray] = None,
        velocity_set=None,
        precision_policy=None,
        compute_backend=None,
        indices=None,
        mesh_vertices=None,
    ):
        assert bc_type in ["velocity", "pressure"], f"type = {bc_type} not supported! Use 'pressure' or 'velocity'."
        self.bc_type = bc_type
        self.native_kind = self._kind_by_type[bc_type]
        self.equilibrium_operator = QuadraticEquilibrium()
        self.profile = profile
        super().__init__(ImplementationStep.STREAMING, velocity_set, precision_policy, compute_backend, indices, mesh_vertices)

        self.prescribed_value = None
        if prescribed_value is not None:
            if profile is not None:
                raise ValueError("Cannot specify both profile and prescribed_value")
            # validation as in the reference (bc_zouhe.py:69-101)
            if isinstance(prescribed_value, (tuple, list)):
                prescribed_value = np.array(prescribed_value, dtype=np.float64)
            elif isinstance(prescribed_value, (int, float)):
                if bc_type == "pressure":
                    prescribed_value = float(prescribed_value)
                else:
                    raise ValueError("Velocity prescribed_value must be a tuple or array")
            elif isinstance(prescribed_value, np.ndarray):
                prescribed_value = prescribed_value.astype(np.float64)
            if bc_type == "velocity":
                if not isinstance(prescribed_value, np.ndarray):
                    raise ValueError("Velocity prescribed_value must be an array-like")
                if np.count_nonzero(prescribed_value) > 1:
                    raise ValueError("This BC only supports normal prescribed values (only one non-zero element allowed)")
            self.prescribed_value = prescribed_value

        self.needs_aux_init = True  # prescribed value is written into f_1 before the first step
        self.needs_aux_recovery = True  # and handed from buffer to buffer every step
        self.num_of_aux_data = 1
        self.needs_padding = True

    # one scalar per BC cell, float64; rounded to the store dtype when written (bc_zouhe.py:96-99)
    def _prescribed_values_at(self, cells_global, missing_cells, global_shape):
        n_cells = cells_global.shape[1]
        d = self.velocity_set.d
        if self.profile is None:
            pv = self.prescribed_value
            if pv is None:
                raise ValueError(f"{type(self).__name__}: neither profile nor prescribed_value was given")
            if np.ndim(pv) == 0:
                return np.full(n_cells, float(pv))
            if self.compute_backend == ComputeBackend.WARP or self.bc_type == "pressure":
                nz = np.nonzero(pv)[0]
                return np.full(n_cells, float(pv[nz][0]) if nz.size else 0.0)
            vec = np.asarray(pv, dtype=np.float64).reshape(-1)[:d, None] * np.ones((1, n_cells))
            return self._project_on_normal(vec, missing_cells)
        if self.compute_backend == ComputeBackend.WARP:
            return _evaluate_index_profile(self.profile, cells_global)
        # JAX convention: profile() -> array broadcastable to (d, *face) (velocity) or (*face) / scalar (pressure)
        values = np.asarray(self.profile(), dtype=np.float64)
        full_shape = tuple(global_shape[:d])
        if self.bc_type == "pressure":
            field = _broadcast_like_reference(values.reshape((1,) + values.shape) if values.ndim < d + 1 else values, (1,) + full_shape)
            return field[(0,) + tuple(cells_global[a] for a in range(d))]
        field = _broadcast_like_reference(values, (d,) + full_shape)
        vec = field[(slice(None),) + tuple(cells_global[a] for a in range(d))]
        return self._project_on_normal(vec, missing_cells)

    def _project_on_normal(self, vec, missing_cells):
        normals = HelperFunctionsBC(self.velocity_set, self.precision_policy, self.compute_backend).get_normal_vectors(missing_cells)
        return -(vec * normals).sum(axis=0)

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_jax(f_pre, f_post, bc_mask, missing_mask)

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_pre, f_post, bc_mask, missing_mask):
        return self._call_warp(f_pre, f_post, bc_mask, missing_mask)


def _broadcast_like_reference(values, target_shape):
    """Insert singleton axes after the leading one, then broadcast (reference: bc_zouhe.py:143-183)."""
    values = np.asarray(values)
    if values.ndim == 2 and values.shape[1] == 1 and len(target_shape) > 2:
        values = values[:, 0]
    if values.ndim < len(target_shape):
        values = values.reshape((values.shape[0],) + (1,) * (len(target_shape) - values.ndim) + values.shape[1:]) if values.ndim else values
    return np.broadcast_to(values, target_shape)


class _IndexVector:
    """Stands in for ``wp.vec3i`` when a per-cell profile is evaluated for ALL cells at once: ``index[a]`` is an array."""

    def __init__(self, rows):
        self._rows = rows

    def __getitem__(self, a):
        return self._rows[a]

    def __len__(self):
        return len(self._rows)


def _evaluate_index_profile(profile, cells_global):
    """profile(index) -> length-1 vector (reference scripts: examples/cfd/flow_past_sphere_3d.py:83-97)."""
    n_cells = cells_global.shape[1]
    rows = [cells_global[a] for a in range(cells_global.shape[0])]
    while len(rows) < 3:
        rows.append(np.zeros(n_cells, dtype=np.int64))
    try:  # vectorised: works when the profile only uses array-aware arithmetic (numpy / the bundled `warp` stand-in)
        out = profile(_IndexVector(rows))
        first = np.asarray(out[0] if isinstance(out, (tuple, list)) or (hasattr(out, "__len__") and not isinstance(out, np.ndarray)) else out, dtype=np.float64)
        if isinstance(out, np.ndarray) and out.ndim == 2:
            first = np.asarray(out[0], dtype=np.float64)
        if first.shape == (n_cells,):
            return first
        if first.ndim == 0:
            return np.full(n_cells, float(first))
    except Exception:
        pass
    values = np.empty(n_cells, dtype=np.float64)
    for i in range(n_cells):
        out = profile(tuple(int(r[i]) for r in rows))
        values[i] = float(out[0] if hasattr(out, "__len__") else out)
    return values
