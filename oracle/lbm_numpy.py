"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
(`xlb_b200`) never imports it and has no CPU fallback.

What this is: a plain-numpy, whole-array restatement of the reference's
**JAX pull path** of the fused lattice-Boltzmann step, operator by operator.
Each function cites the reference file:line it follows (paths relative to
/root/reference).  Where the reference's Warp path differs *observably* from
the JAX path the function takes ``flavor="jax"|"warp"`` and the difference is
described in the docstring.

Parity status: PINNED.  (i) The known-answer tests of the reference's own test
suite for this path are reproduced in tests/test_oracle_known_answers.py.
(ii) The reference's own Python (JAX backend) was executed in the build
container under a numpy-backed ``jax`` stand-in (oracle/refshim; real jax is not
installable here) to generate tests/golden/*.npz with tests/golden/make_golden.py;
tests/test_oracle_vs_golden.py checks this oracle against those vectors.

Array layout everywhere: ``f[q, nx, ny(, nz)]`` exactly as the reference
(`grid.create_field`, xlb/grid/warp_grid.py:17-32).
"""

from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

# --------------------------------------------------------------------------------------------
# Lattices  (xlb/velocity_set/velocity_set.py:55-221, d3q19.py:19-27, d3q27.py:19-29, d2q9.py:18-21)
# --------------------------------------------------------------------------------------------


class Lattice:
    def __init__(self, name: str):
        name = name.upper()
        if name == "D3Q19":
            c = [v for v in itertools.product([0, -1, 1], repeat=3) if abs(v[0]) + abs(v[1]) + abs(v[2]) <= 2]
            wt = {0: 1 / 3, 1: 1 / 18, 2: 1 / 36}
        elif name == "D3Q27":
            c = list(itertools.product([0, -1, 1], repeat=3))
            wt = {0: 8 / 27, 1: 2 / 27, 2: 1 / 54, 3: 1 / 216}
        elif name == "D2Q9":
            c = list(zip([0, 0, 0, 1, -1, 1, -1, 1, -1], [0, 1, -1, 0, 1, -1, 0, 1, -1]))
            wt = {0: 4 / 9, 1: 1 / 9, 2: 1 / 36}
        else:
            raise ValueError(name)
        self.name = name
        self.c = np.array(c, dtype=np.int64).T  # (d, q)
        self.d, self.q = self.c.shape
        self.w = np.array([wt[int(np.abs(v).sum())] for v in self.c.T])
        cl = self.c.T.tolist()
        self.opp = np.array([cl.index([-a for a in v]) for v in cl])
        nt = self.d * (self.d + 1) // 2
        self.cc = np.zeros((self.q, nt))
        t = 0
        for a in range(self.d):
            for b in range(a, self.d):
                self.cc[:, t] = self.c[a] * self.c[b]
                t += 1
        self.qi = self.cc.copy()
        diag, off = ((0, 3, 5), (1, 2, 4)) if self.d == 3 else ((0, 2), (1,))
        self.qi[:, diag] += -1.0 / 3.0
        self.qi[:, off] *= 2.0
        speed = np.abs(self.c).sum(axis=0)
        self.main = np.nonzero(speed == 1)[0]
        self.right = np.nonzero(self.c[0] == 1)[0]
        self.left = np.nonzero(self.c[0] == -1)[0]


# --------------------------------------------------------------------------------------------
# Precision policy  (xlb/precision_policy.py:46-89)
# --------------------------------------------------------------------------------------------

_NP = {"FP64": np.float64, "FP32": np.float32, "FP16": np.float16}


def policy_dtypes(policy: str):
    """'FP32FP16' -> (compute dtype, store dtype)."""
    return _NP[policy[:4]], _NP[policy[4:]]


# --------------------------------------------------------------------------------------------
# Element operators
# --------------------------------------------------------------------------------------------


def stream(f: np.ndarray, lat: Lattice) -> np.ndarray:
    """Pull streaming with periodic wrap: out[l, x] = f[l, x - c_l]  (stream.py:18-51)."""
    out = np.empty_like(f)
    axes = tuple(range(lat.d))
    for l in range(lat.q):
        out[l] = np.roll(f[l], tuple(int(s) for s in lat.c[:, l]), axis=axes)
    return out


def zero_moment(f):
    """rho = sum_l f_l, keepdims (zero_moment.py:14-17)."""
    return np.sum(f, axis=0, keepdims=True)


def first_moment(f, rho, lat: Lattice):
    """u = (c . f) / rho  (first_moment.py:14-18)."""
    c = lat.c.astype(f.dtype)
    return np.tensordot(c, f, axes=(-1, 0)) / rho


def macroscopic(f, lat: Lattice):
    """(rho, u)  (macroscopic.py:26-31)."""
    rho = zero_moment(f)
    return rho, first_moment(f, rho, lat)


def second_moment(f, lat: Lattice):
    """Pi_t = sum_l cc[l,t] f_l, t over xx,xy,xz,yy,yz,zz  (second_moment.py:35-55)."""
    return np.tensordot(lat.cc.astype(f.dtype), f, axes=(0, 0))


def equilibrium(rho, u, lat: Lattice):
    """feq = rho w (1 + cu(1 + 0.5 cu) - usqr), cu = 3 c.u, usqr = 1.5 u.u  (quadratic_equilibrium.py:18-25)."""
    dt = u.dtype
    cu = dt.type(3.0) * np.tensordot(lat.c.astype(dt), u, axes=(0, 0))
    usqr = dt.type(1.5) * np.sum(np.square(u), axis=0, keepdims=True)
    w = lat.w.astype(dt).reshape((-1,) + (1,) * (u.ndim - 1))
    return rho * w * (dt.type(1.0) + cu * (dt.type(1.0) + dt.type(0.5) * cu) - usqr)


def collide_bgk(f, feq, omega):
    """f - omega (f - feq), omega cast to the compute dtype  (bgk.py:17-22)."""
    return f - f.dtype.type(omega) * (f - feq)


def _kbc_shear(fneq, lat: Lattice):
    """Shear part of fneq  (kbc.py:102-180)."""
    pi = second_moment(fneq, lat)
    s = np.zeros_like(fneq)
    if lat.name == "D3Q27":
        nxz = pi[0] - pi[5]
        nyz = pi[3] - pi[5]
        s[9] = s[18] = (2.0 * nxz - nyz) / 6.0
        s[3] = s[6] = (-nxz + 2.0 * nyz) / 6.0
        s[1] = s[2] = (-nxz - nyz) / 6.0
        s[12] = s[24] = pi[1] / 4.0
        s[21] = s[15] = -pi[1] / 4.0
        s[10] = s[20] = pi[2] / 4.0
        s[19] = s[11] = -pi[2] / 4.0
        s[8] = s[4] = pi[4] / 4.0
        s[7] = s[5] = -pi[4] / 4.0
    elif lat.name == "D2Q9":
        n = pi[0] - pi[2]
        s[3] = s[6] = n
        s[1] = s[2] = -n
        s[7] = s[8] = pi[1]
        s[4] = s[5] = -pi[1]
    else:
        raise NotImplementedError("KBC: velocity set not supported: " + lat.name)  # kbc.py:71-72
    return s


def collide_kbc(f, feq, rho, lat: Lattice, omega, epsilon=1e-32):
    """KBC entropic-stabiliser collision  (kbc.py:40-100)."""
    dt = f.dtype.type
    fneq = f - feq
    shear = _kbc_shear(fneq, lat)
    delta_s = shear * rho / dt(4.0) if lat.d == 2 else shear * rho
    beta = dt(0.5) * dt(omega)
    inv_beta = dt(1.0) / beta
    delta_h = fneq - delta_s
    temp = delta_h / feq
    sp1 = np.sum(temp * delta_s, axis=0)
    sp2 = np.sum(temp * delta_h, axis=0)
    gamma = inv_beta - (dt(2.0) - inv_beta) * sp1 / (dt(epsilon) + sp2)
    return f - beta * (dt(2.0) * delta_s + gamma[None, ...] * delta_h)


def collide_smagorinsky(f, feq, lat: Lattice, omega, coef=0.17):
    """SmagorinskyLESBGK (smagorinsky_les_bgk.py:37-90; the reference has a Warp functional only, and it indexes c[2, l],
    i.e. 3-D only).  The 'strain' it uses is a weighted sum of squared non-equilibrium POPULATIONS selected by the SIGNED
    component sum of c_l (== 1: weight 1, >= 2: weight 2), accumulated in l order; restated literally."""
    if lat.d != 3:
        raise ValueError("SmagorinskyLESBGK: the reference functional reads c[2, l] (3-D lattices only)")
    dt = f.dtype.type
    fneq = f - feq
    csum = lat.c.sum(axis=0)
    strain = np.zeros(f.shape[1:], dtype=f.dtype)
    for l in range(lat.q):
        if csum[l] == 1:
            strain = strain + fneq[l] * fneq[l]
        if csum[l] >= 2:
            strain = strain + dt(2.0) * fneq[l] * fneq[l]
    tau0 = dt(1.0) / dt(omega)
    tau = tau0 + dt(0.5) * (np.sqrt(tau0 * tau0 + dt(36.0) * (dt(coef) ** dt(2.0)) * np.sqrt(strain)) - tau0)
    return f - (dt(1.0) / tau)[None, ...] * fneq


# --------------------------------------------------------------------------------------------
# Boundary conditions
# --------------------------------------------------------------------------------------------

STREAMING, COLLISION = "streaming", "collision"

_KIND_STEP = {
    "equilibrium": STREAMING,
    "donothing": STREAMING,
    "halfway": STREAMING,
    "fullway": COLLISION,
    "zouhe": STREAMING,
    "regularized": STREAMING,
    "outflow": STREAMING,
}
_NEEDS_PADDING = {"halfway", "zouhe", "regularized"}  # bc_halfway_bounce_back.py:48, bc_zouhe.py:115


@dataclass
class BC:
    """Plain description of one boundary condition (stands in for the reference's BC objects).

    kind: equilibrium | donothing | halfway | fullway | zouhe | regularized | outflow
    id:   uint8 id written into bc_mask (registry order in the reference, boundary_condition_registry.py:19-27)
    indices: (d, n) integer array, global cell coordinates
    rho, u: EquilibriumBC parameters (bc_equilibrium.py:42-44)
    bc_type: 'velocity' | 'pressure' for zouhe / regularized (bc_zouhe.py:52)
    prescribed: zouhe/regularized prescribed values.  JAX convention (bc_zouhe.py:120-121, 130-199):
        velocity -> array broadcastable to (d, *grid) [a (d,) vector, or (d, ny, nz) profile]; pressure -> scalar/array.
    """

    kind: str
    id: int
    indices: np.ndarray
    rho: float = 1.0
    u: Sequence[float] = (0.0, 0.0, 0.0)
    bc_type: str = "velocity"
    prescribed: Optional[np.ndarray] = None
    normal: Optional[np.ndarray] = field(default=None)

    def __post_init__(self):
        self.indices = np.asarray(self.indices, dtype=np.int64)
        self.step = _KIND_STEP[self.kind]
        self.needs_padding = self.kind in _NEEDS_PADDING
        if self.kind == "outflow" and self.normal is None:
            self.normal = outflow_normal(self.indices)


def outflow_normal(indices) -> np.ndarray:
    """Outward normal of a flat axis-aligned face from index statistics (bc_extrapolation_outflow.py:63-77)."""
    from collections import Counter

    freq = [Counter(np.asarray(coord).tolist()).most_common(1)[0] for coord in indices]
    counts = np.array([cnt for _, cnt in freq])
    elements = np.array([el for el, _ in freq])
    normal = counts // counts.max()
    if elements[np.argmax(counts)] == 0:
        normal = normal * -1
    return normal


def _bmask(bc_mask, bc_id, q):
    """(bc_mask == id) broadcast over q  (e.g. bc_fullway_bounce_back.py:47-49)."""
    b = bc_mask == bc_id
    return np.broadcast_to(b, (q,) + b.shape[1:])


def bc_equilibrium(bc: BC, f_pre, f_post, bc_mask, missing, lat):
    """f = feq(rho_bc, u_bc) on the BC cells  (bc_equilibrium.py:58-66)."""
    dt = f_post.dtype
    feq = equilibrium(np.array([bc.rho], dtype=dt), np.array(bc.u[: lat.d], dtype=dt), lat)  # shape (q,)
    feq = feq.reshape((lat.q,) + (1,) * lat.d)
    return np.where(bc_mask == bc.id, feq, f_post)


def bc_donothing(bc, f_pre, f_post, bc_mask, missing, lat):
    """f = f_pre on the BC cells  (bc_do_nothing.py:44-48)."""
    return np.where(bc_mask == bc.id, f_pre, f_post)


def bc_fullway(bc, f_pre, f_post, bc_mask, missing, lat):
    """f[l] = f_pre[opp[l]] on the BC cells  (bc_fullway_bounce_back.py:44-50)."""
    return np.where(_bmask(bc_mask, bc.id, lat.q), f_pre[lat.opp], f_post)


def bc_halfway(bc, f_pre, f_post, bc_mask, missing, lat):
    """missing l on BC cells: f[l] = f_pre[opp[l]]  (bc_halfway_bounce_back.py:50-60)."""
    return np.where(np.logical_and(missing, _bmask(bc_mask, bc.id, lat.q)), f_pre[lat.opp], f_post)


def _normals(missing, lat, flavor):
    """Outward normal n = -c_l of the missing axis-aligned direction(s).

    jax: minus the SUM over all missing main directions (bc_zouhe.py:137-141);
    warp: minus the FIRST missing main direction in index order (helper_functions_bc.py:75-86).
    Identical on flat faces (one missing main direction)."""
    main_c = lat.c[:, lat.main]  # (d, nmain)
    m = missing[lat.main]
    if flavor == "jax":
        return -np.tensordot(main_c, m.astype(np.int64), axes=(-1, 0))
    n = np.zeros((lat.d,) + missing.shape[1:], dtype=np.int64)
    found = np.zeros(missing.shape[1:], dtype=bool)
    for j, l in enumerate(lat.main):
        sel = m[j] & ~found
        for a in range(lat.d):
            n[a][sel] = -lat.c[a, l]
        found |= sel
    return n


def _fsum(fpop, missing, lat):
    """sum_middle f + 2 sum_known f; known = missing[opp], middle = neither  (bc_zouhe.py:131-134, 212-213)."""
    known = missing[lat.opp]
    middle = ~(missing | known)
    return np.sum(fpop * middle, axis=0, keepdims=True) + fpop.dtype.type(2.0) * np.sum(fpop * known, axis=0, keepdims=True)


def _broadcast_prescribed(p, target_shape):
    """(d,) / (d,1) vectors and (d, ny, nz) profiles -> (d, *grid)  (bc_zouhe.py:143-183)."""
    p = np.asarray(p)
    if p.ndim == 2 and p.shape[1] == 1:
        p = p[:, 0]
    if p.ndim < len(target_shape):
        p = p.reshape((p.shape[0],) + (1,) * (len(target_shape) - p.ndim) + p.shape[1:]) if p.ndim > 0 else p.reshape((1,) * len(target_shape))
    return np.broadcast_to(p, target_shape)


def _zouhe_equilibrium(bc: BC, f_post, missing, lat, flavor):
    """rho, u at the boundary then feq  (bc_zouhe.py:185-243 JAX; 279-343 Warp)."""
    dt = f_post.dtype
    normals = _normals(missing, lat, flavor).astype(dt)
    fsum = _fsum(f_post, missing, lat)
    if bc.bc_type == "velocity":
        if flavor == "jax":
            vel = _broadcast_prescribed(np.asarray(bc.prescribed, dtype=dt), (lat.d,) + f_post.shape[1:])
        else:
            # Warp keeps ONE scalar per cell = magnitude of the normal velocity and rebuilds u = -value * n
            # (bc_zouhe.py:96-99, 302-303).  `prescribed` is then a scalar or a (*grid) array.
            vel = -np.asarray(bc.prescribed, dtype=dt) * normals
        unormal = np.sum(normals * vel, keepdims=True, axis=0)
        rho = fsum / (dt.type(1.0) + unormal)
    else:
        rho = np.asarray(bc.prescribed, dtype=dt)
        unormal = dt.type(-1.0) + fsum / rho
        vel = unormal * normals
        rho = np.broadcast_to(rho, fsum.shape)
    return equilibrium(rho, vel, lat)


def _bounceback_nonequilibrium(fpop, feq, missing, lat):
    """missing l: f[l] = f[opp] + feq[l] - feq[opp]  (bc_zouhe.py:245-254)."""
    fknown = fpop[lat.opp] + feq - feq[lat.opp]
    return np.where(missing, fknown, fpop)


def _regularize(fpop, feq, lat):
    """f = feq + 4.5 w (Qi : Pi_neq)  (bc_regularized.py:67-105)."""
    dt = fpop.dtype
    w = lat.w.astype(dt).reshape((-1,) + (1,) * lat.d)
    pineq = second_moment(fpop - feq, lat)
    qipi = np.tensordot(lat.qi.astype(dt), pineq, axes=(1, 0))
    return feq + dt.type(9.0 / 2.0) * w * qipi


def bc_zouhe(bc, f_pre, f_post, bc_mask, missing, lat, flavor="jax"):
    """Zou-He: non-equilibrium bounce-back of the unknown populations  (bc_zouhe.py:256-269)."""
    with np.errstate(all="ignore"):  # non-BC cells divide by garbage and are masked out
        feq = _zouhe_equilibrium(bc, f_post, missing, lat, flavor)
        f_bd = _bounceback_nonequilibrium(f_post, feq, missing, lat)
    return np.where(_bmask(bc_mask, bc.id, lat.q), f_bd, f_post)


def bc_regularized(bc, f_pre, f_post, bc_mask, missing, lat, flavor="jax"):
    """Zou-He followed by regularisation of ALL populations  (bc_regularized.py:107-124)."""
    with np.errstate(all="ignore"):
        feq = _zouhe_equilibrium(bc, f_post, missing, lat, flavor)
        f_bd = _bounceback_nonequilibrium(f_post, feq, missing, lat)
        f_bd = _regularize(f_bd, feq, lat)
    return np.where(_bmask(bc_mask, bc.id, lat.q), f_bd, f_post)


def bc_outflow(bc, f_pre, f_post, bc_mask, missing, lat):
    """Streaming part: missing l: f[l] = f_pre[opp[l]] (aux stored last step)  (bc_extrapolation_outflow.py:120-129)."""
    return np.where(np.logical_and(missing, _bmask(bc_mask, bc.id, lat.q)), f_pre[lat.opp], f_post)


def outflow_update_aux(bc, f_pre, f_post, bc_mask, missing, lat):
    """Post-collision aux update  (bc_extrapolation_outflow.py:91-118).

    f_pre = post-stream (post-BC) populations, f_post = post-collision.  For outlet cells and directions l that are
    `known` (missing[opp[l]]): f_post[l] = cs * f_pre[opp[l]](cell - n) + (1 - cs) * f_pre[opp[l]](cell), cs = 1/sqrt(3).
    The Warp path computes the same value by pulling f_0[l', cell - (c_l' + n)] directly
    (bc_extrapolation_outflow.py:175-194), which equals the neighbour's post-stream value when the neighbour has no BC.
    """
    dt = f_post.dtype
    cs = dt.type(1.0) / np.sqrt(dt.type(3.0))  # jax: 1.0 / jnp.sqrt(3.0) evaluated in the compute dtype
    boundary = _bmask(bc_mask, bc.id, lat.q)
    nrm = tuple(int(v) for v in bc.normal[: lat.d])
    axes = tuple(range(1, lat.d + 1))
    neighbour = np.roll(boundary, tuple(-v for v in nrm), axis=axes)
    fpop = np.where(boundary, f_pre, f_post)
    fpop_nb = np.where(neighbour, f_pre, f_post)
    fpop_nb = np.roll(fpop_nb, nrm, axis=axes)
    extrap = cs * fpop_nb + (dt.type(1.0) - cs) * fpop
    known = missing[lat.opp]
    return np.where(np.logical_and(boundary, known), extrap[lat.opp], f_post)


_APPLY = {
    "equilibrium": bc_equilibrium,
    "donothing": bc_donothing,
    "fullway": bc_fullway,
    "halfway": bc_halfway,
    "outflow": bc_outflow,
}


def apply_bc(bc: BC, f_pre, f_post, bc_mask, missing, lat, flavor="jax"):
    if bc.kind == "zouhe":
        return bc_zouhe(bc, f_pre, f_post, bc_mask, missing, lat, flavor)
    if bc.kind == "regularized":
        return bc_regularized(bc, f_pre, f_post, bc_mask, missing, lat, flavor)
    return _APPLY[bc.kind](bc, f_pre, f_post, bc_mask, missing, lat)


# --------------------------------------------------------------------------------------------
# Boundary masks  (indices_boundary_masker.py)
# --------------------------------------------------------------------------------------------


def indices_in_interior(indices, shape):
    """Strictly interior test per index  (indices_boundary_masker.py:29-40)."""
    d = len(shape)
    sh = np.array(shape)
    return np.all((indices[:d] > 0) & (indices[:d] < sh[:d, None] - 1), axis=0)


def build_masks_jax(bcs: Sequence[BC], shape, lat: Lattice, n_devices: int = 1):
    """JAX masker  (indices_boundary_masker.py:45-101).

    Domain padded (x by n_devices, y/z by 1) with 'solid'; ids written per BC in list order; BCs with needs_padding and
    ANY strictly interior index mark their cells solid and push their id to all neighbours; finally
    missing[l, x] = solid(x - c_l) for EVERY cell."""
    d = lat.d
    pad = (n_devices,) + (1,) * (d - 1)
    pshape = tuple(s + 2 * p for s, p in zip(shape, pad))
    bmap = np.zeros(pshape, dtype=np.uint8)
    solid = np.ones((lat.q,) + pshape, dtype=bool)
    inner = tuple(slice(p, -p) for p in pad)
    solid[(slice(None),) + inner] = False
    for bc in bcs:
        idx = bc.indices
        pidx = idx + np.array(pad)[:, None]
        bmap[tuple(pidx)] = bc.id
        if bc.needs_padding and np.any(indices_in_interior(idx, shape)):
            solid[(slice(None),) + tuple(pidx)] = True
            push = (pidx[:, :, None] + lat.c[:, None, :]).reshape(d, -1)
            bmap[tuple(push)] = bc.id
    missing = stream(solid, lat)[(slice(None),) + inner]
    return bmap[inner][None].copy(), missing.copy()


def build_masks_warp(bcs: Sequence[BC], shape, lat: Lattice):
    """Warp masker  (indices_boundary_masker.py:103-224).

    Per index: bc_mask[idx] = id; missing[l, idx] = True if idx - c_l is outside the domain; otherwise, if that index
    is strictly interior (per-index flag, only for needs_padding BCs): missing[l, idx + c_l] = True and
    bc_mask[idx + c_l] = id.  BCs are processed in list order here (the reference launches them in one kernel and is
    last-writer-wins on overlap, which `check_bc_overlaps` forbids)."""
    d = lat.d
    sh = np.array(shape)[:, None]
    bc_mask = np.zeros((1,) + tuple(shape), dtype=np.uint8)
    missing = np.zeros((lat.q,) + tuple(shape), dtype=bool)
    for bc in bcs:
        idx = bc.indices
        inb = np.all((idx >= 0) & (idx < sh), axis=0)
        idx = idx[:, inb]
        interior = indices_in_interior(idx, shape) if bc.needs_padding else np.zeros(idx.shape[1], dtype=bool)
        bc_mask[(0,) + tuple(idx)] = bc.id
        for l in range(lat.q):
            pull = idx - lat.c[:, l : l + 1]
            oob = ~np.all((pull >= 0) & (pull < sh), axis=0)
            missing[(l,) + tuple(idx[:, oob])] = True
            sel = (~oob) & interior
            push = idx[:, sel] + lat.c[:, l : l + 1]
            missing[(l,) + tuple(push)] = True
            bc_mask[(0,) + tuple(push)] = bc.id
    return bc_mask, missing


def build_masks(bcs, shape, lat, flavor="warp", n_devices=1):
    return build_masks_jax(bcs, shape, lat, n_devices) if flavor == "jax" else build_masks_warp(bcs, shape, lat)


def bounding_box_indices(shape, remove_edges=False):
    """Face index arrays  (grid/grid.py:34-86)."""
    shape = tuple(shape)
    d = len(shape)
    grid = np.indices(shape)
    o = 1 if remove_edges else 0
    sl = [slice(o, n - o) for n in shape]
    if d == 2:
        nx, ny = shape
        out = {"bottom": grid[:, sl[0], 0], "top": grid[:, sl[0], ny - 1], "left": grid[:, 0, sl[1]], "right": grid[:, nx - 1, sl[1]]}
    else:
        nx, ny, nz = shape
        out = {
            "bottom": grid[:, sl[0], sl[1], 0],
            "top": grid[:, sl[0], sl[1], nz - 1],
            "left": grid[:, 0, sl[1], sl[2]],
            "right": grid[:, nx - 1, sl[1], sl[2]],
            "front": grid[:, sl[0], 0, sl[2]],
            "back": grid[:, sl[0], ny - 1, sl[2]],
        }
    return {k: v.reshape(d, -1) for k, v in out.items()}


# --------------------------------------------------------------------------------------------
# The step  (nse_stepper.py:147-192, JAX pull scheme)
# --------------------------------------------------------------------------------------------


def exact_difference_force(f_post_collision, feq, rho, u, force, lat: Lattice):
    """ExactDifference body force: f += feq(rho, u + F) - feq(rho, u)  (force/exact_difference_force.py:45-70;
    applied after the collision by ForcedCollision, collision/forced_collision.py:34-39).  Not built in xlb_b200 yet
    (SURVEY.md §8f N4): the oracle and its golden vector are the target for the next round."""
    dt = u.dtype
    delta_u = np.asarray(force, dtype=dt).reshape((lat.d,) + (1,) * lat.d)
    return f_post_collision + (equilibrium(rho, u + delta_u, lat) - feq)


def step(f0, bc_mask, missing, bcs: Sequence[BC], omega, lat: Lattice, policy="FP32FP32", collision="BGK", flavor="jax", f1_prev=None, force=None,
         smagorinsky=0.17):
    """One pull step.  Returns f1 in the store dtype.

    flavor="warp" adds the two observable Warp-only behaviours: cells with bc_mask == 255 are skipped entirely
    (nse_stepper.py:356-358; they keep `f1_prev`), and Zou-He/Regularized use the scalar-normal convention."""
    cdt, sdt = policy_dtypes(policy)
    f = f0.astype(cdt)  # cast_to_compute (nse_stepper.py:153)
    f_post = stream(f, lat)  # L157
    for bc in bcs:  # L160-167
        if bc.step == STREAMING:
            f_post = apply_bc(bc, f, f_post, bc_mask, missing, lat, flavor)
    rho, u = macroscopic(f_post, lat)  # L170
    feq = equilibrium(rho, u, lat)  # L173
    if collision == "BGK":  # L176
        f_out = collide_bgk(f_post, feq, omega)
    elif collision == "KBC":
        with np.errstate(all="ignore"):
            f_out = collide_kbc(f_post, feq, rho, lat, omega)
    elif collision == "SmagorinskyLESBGK":  # nse_stepper.py:42-43
        f_out = collide_smagorinsky(f_post, feq, lat, omega, smagorinsky)
    else:
        raise ValueError(collision)
    if force is not None:
        f_out = exact_difference_force(f_out, feq, rho, u, force, lat)
    for bc in bcs:  # L179-187
        if bc.kind == "outflow":
            f_out = outflow_update_aux(bc, f_post, f_out, bc_mask, missing, lat)
        if bc.step == COLLISION:
            f_out = apply_bc(bc, f_post, f_out, bc_mask, missing, lat, flavor)
    f1 = f_out.astype(sdt)  # cast_to_store (L190)
    if flavor == "warp" and np.any(bc_mask == 255):
        keep = f1_prev if f1_prev is not None else f0
        f1 = np.where(bc_mask == 255, keep.astype(sdt), f1)
    return f1


def run(f0, bc_mask, missing, bcs, omega, lat, nsteps, policy="FP32FP32", collision="BGK", flavor="jax", force=None, smagorinsky=0.17):
    """The user loop of examples/performance/mlups_3d.py:77-80: step then swap."""
    f_a, f_b = f0, f0.copy()
    for _ in range(nsteps):
        f_b = step(f_a, bc_mask, missing, bcs, omega, lat, policy, collision, flavor, f1_prev=f_b, force=force, smagorinsky=smagorinsky)
        f_a, f_b = f_b, f_a
    return f_a


def initialize_eq(shape, lat: Lattice, policy="FP32FP32", rho=None, u=None):
    """f = feq(rho, u) in the store dtype; default rho = 1, u = 0  (helper/initializers.py:5-20)."""
    cdt, sdt = policy_dtypes(policy)
    rho = np.ones((1,) + tuple(shape), dtype=cdt) if rho is None else np.asarray(rho, dtype=cdt)
    u = np.zeros((lat.d,) + tuple(shape), dtype=cdt) if u is None else np.asarray(u, dtype=cdt)
    return equilibrium(rho, u, lat).astype(sdt)


def momentum_transfer(bc: BC, f_post_collision, bc_mask, missing, lat: Lattice, flavor="jax"):
    """Net momentum-exchange force on the solid behind the no-slip BC  (force/momentum_transfer.py:51-90)."""
    f_pc = f_post_collision
    f_ps = apply_bc(bc, f_pc, stream(f_pc, lat), bc_mask, missing, lat, flavor)
    boundary = _bmask(bc_mask, bc.id, lat.q)
    is_edge = np.logical_and(boundary, ~missing[0])
    phi = f_pc[lat.opp] + f_ps
    phi = np.where(np.logical_and(missing, is_edge), phi, f_pc.dtype.type(0.0))
    force = np.tensordot(lat.c[:, lat.opp].astype(f_pc.dtype), phi, axes=(-1, 0))
    return force.sum(axis=tuple(range(1, lat.d + 1)))


# --------------------------------------------------------------------------------------------
# Mesh-based masks  (boundary_masker/mesh_boundary_masker.py:49-236)
# --------------------------------------------------------------------------------------------


def _tri_setup(v0, v1, v2, edge_test):
    """pre_compute (mesh_boundary_masker.py:65-99) in float32, one rounding per operation in the reference's order.
    edge_test="reference": the edge normals / offsets exactly as written there (both components from e[ax0], v[ax0]);
    edge_test="schwarz_seidel": the published form the reference cites (n_e = sgn (-e[ax1], e[ax0]), offset from v[ax0], v[ax1])."""
    f = np.float32
    v = [np.asarray(x, dtype=f) for x in (v0, v1, v2)]
    n = np.cross(v[1] - v[0], v[2] - v[0]).astype(f)
    length = np.sqrt((n * n).sum(dtype=f))
    valid = bool(length > 0)
    n = (n / length).astype(f) if valid else np.zeros(3, f)
    corner = (n > 0).astype(f)
    dist1 = (n * (corner - v[0])).sum(dtype=f)
    dist2 = (n * ((f(1.0) - corner) - v[0])).sum(dtype=f)
    e = [v[(i + 1) % 3] - v[i] for i in range(3)]
    ne0, ne1, de = np.zeros((3, 3), f), np.zeros((3, 3), f), np.zeros((3, 3), f)
    for ax0 in range(3):
        ax1, ax2 = (ax0 + 1) % 3, (ax0 + 2) % 3
        sgn = f(-1.0) if n[ax2] < 0 else f(1.0)
        for i in range(3):
            a, b = (ax0, ax0) if edge_test == "reference" else (ax1, ax1)
            ne0[i, ax0] = f(-1.0) * sgn * e[i][a]
            ne1[i, ax0] = sgn * e[i][ax0]
            vb = v[i][ax0] if edge_test == "reference" else v[i][b]
            de[i, ax0] = (f(-1.0) * (ne0[i, ax0] * v[i][ax0] + ne1[i, ax0] * vb) + max(f(0.0), ne0[i, ax0])) + max(f(0.0), ne1[i, ax0])
    lo = np.minimum(np.minimum(v[0], v[1]), v[2])
    hi = np.maximum(np.maximum(v[0], v[1]), v[2])
    return dict(n=n, dist1=dist1, dist2=dist2, ne0=ne0, ne1=ne1, de=de, lo=lo, hi=hi, valid=valid)


def _tri_box_overlap(t, low):
    """triangle_box_intersect (L110-128) for unit boxes at `low` [..., 3] (float32), after the inclusive bounding-box
    overlap that wp.mesh_query_aabb applies (L136)."""
    f = np.float32
    low = np.asarray(low, dtype=f)
    hit = np.all((t["lo"] <= low + f(1.0)) & (t["hi"] >= low), axis=-1)
    if not t["valid"]:
        return np.zeros_like(hit)
    n = t["n"]
    nl = (n[0] * low[..., 0] + n[1] * low[..., 1]) + n[2] * low[..., 2]
    hit &= (nl + t["dist1"]) * (nl + t["dist2"]) <= 0
    for ax0 in range(3):
        ax1 = (ax0 + 1) % 3
        for i in range(3):
            hit &= (t["ne0"][i, ax0] * low[..., ax0] + t["ne1"][i, ax0] * low[..., ax1]) + t["de"][i, ax0] >= 0
    return hit


def mesh_solid_voxels(vertices, shape, edge_test="schwarz_seidel"):
    """bool [nx+2, ny+2, nz+2]: voxel (i, j, k) -> [i+1, j+1, k+1] is True iff some triangle overlaps the box [i, i+1]^3
    (mesh_voxel_intersect, L133-148).  vertices: (3 T, 3), three consecutive rows per triangle (L219-223)."""
    tri = np.asarray(vertices, dtype=np.float32).reshape(-1, 3, 3)
    solid = np.zeros(tuple(s + 2 for s in shape), dtype=bool)
    for v0, v1, v2 in tri:
        t = _tri_setup(v0, v1, v2, edge_test)
        lo = np.maximum(-1, np.ceil(t["lo"] - np.float32(1.0)).astype(np.int64))
        hi = np.minimum(np.array(shape), np.floor(t["hi"]).astype(np.int64))
        if np.any(lo > hi):
            continue
        ax = [np.arange(lo[k], hi[k] + 1) for k in range(3)]
        I, J, K = np.meshgrid(*ax, indexing="ij")
        hit = _tri_box_overlap(t, np.stack([I, J, K], axis=-1).astype(np.float32))
        solid[I[hit] + 1, J[hit] + 1, K[hit] + 1] = True
    return solid


def build_masks_mesh(vertices, bc_id, bc_mask, missing, lat: Lattice, edge_test="schwarz_seidel"):
    """MeshBoundaryMasker kernel (L153-190) applied on top of existing masks: solid voxels -> 255; every other cell with a
    solid neighbour in direction l -> bc_mask = id, missing[opp[l]] = True."""
    if lat.d != 3:
        raise NotImplementedError("This Operator is not implemented in 2D!")  # L27-28
    shape = bc_mask.shape[1:]
    solid = mesh_solid_voxels(vertices, shape, edge_test)
    inner = solid[1:-1, 1:-1, 1:-1]
    bc_mask, missing = bc_mask.copy(), missing.copy()
    bc_mask[0][inner] = 255
    for l in range(1, lat.q):
        c = lat.c[:, l]
        nb = solid[1 + c[0] : 1 + c[0] + shape[0], 1 + c[1] : 1 + c[1] + shape[1], 1 + c[2] : 1 + c[2] + shape[2]]
        sel = nb & ~inner
        bc_mask[0][sel] = bc_id
        missing[lat.opp[l]][sel] = True
    return bc_mask, missing


# --------------------------------------------------------------------------------------------
# x-slab halo exchange emulation  (distribute/distribute.py:23-44)
# --------------------------------------------------------------------------------------------


def stream_sharded(f, lat: Lattice, n_shards: int):
    """Each shard streams with LOCAL periodic roll, then the wrongly wrapped planes are swapped ring-wise:
    populations with c_x = +1 on local plane 0 come from the left neighbour's roll result plane 0 ... exactly as
    `lax.ppermute` does in the reference.  Must equal `stream(f)` on the whole domain."""
    shards = np.split(f, n_shards, axis=1)
    rolled = [stream(s, lat) for s in shards]
    out = [r.copy() for r in rolled]
    for i in range(n_shards):
        # right-moving populations that wrapped locally onto plane 0 of shard i belong to shard i+1
        out[(i + 1) % n_shards][lat.right, :1] = rolled[i][lat.right, :1]
        out[(i - 1) % n_shards][lat.left, -1:] = rolled[i][lat.left, -1:]
    return np.concatenate(out, axis=1)
