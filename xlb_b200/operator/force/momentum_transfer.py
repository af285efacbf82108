"""MomentumTransfer: net force on a solid by the momentum-exchange method (Ladd 1994; Mei et al. 2002).

Reference: xlb/operator/force/momentum_transfer.py — JAX L51-90, Warp kernel L108-160 (atomic_add into one vector) and
its launch L162-176.  For every EDGE cell of the no-slip BC (bc id matches and the rest direction is not missing) the
post-collision populations `f_0` are pulled once, the BC's functional is applied, and
``m_d = sum_{missing l} c[d, opp l] (f_0[opp l] + f_post_stream[l])`` is accumulated over all such cells.
Call after the boundary conditions were imposed: ``force = momentum_transfer(f_0, f_1, bc_mask, missing_mask)``.
Native: xlbn_momentum_transfer (accumulated in fp64, one atomic per warp and component).
"""

import torch

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.field import Field
from xlb_b200.operator.operator import Operator
from xlb_b200.operator.stream import Stream


class MomentumTransfer(Operator):
    def __init__(self, no_slip_bc_instance, velocity_set=None, precision_policy=None, compute_backend=None):
        self.no_slip_bc_instance = no_slip_bc_instance
        super().__init__(velocity_set, precision_policy, compute_backend)
        self.stream = Stream(self.velocity_set, self.precision_policy, self.compute_backend)

    def _run(self, f_0, f_1, bc_mask, missing_mask):
        vs = self.velocity_set
        for name, t in (("f_0", f_0), ("bc_mask", bc_mask), ("missing_mask", missing_mask)):
            native.require_cuda(t, name)
        if f_0.shape[0] != vs.q or bc_mask.dtype != torch.uint8 or missing_mask.dtype != torch.bool:
            raise ValueError("MomentumTransfer: f_0 must be [q, ...], bc_mask uint8, missing_mask bool")
        if f_1 is not None and (f_1.shape != f_0.shape or f_1.dtype != f_0.dtype):
            f_1 = None
        dims = native.dims_of(f_0, vs.d)
        force = torch.zeros(3, dtype=torch.float64, device=f_0.device)
        desc = self.no_slip_bc_instance.native_desc()
        native.check(
            native.lib().xlbn_momentum_transfer(
                self._lattice, self._compute_code, desc, native.ptr(f_0), native.ptr(f_1), native.dtype_code(f_0.dtype), native.ptr(bc_mask),
                native.ptr(missing_mask), native.int3(dims), native.ptr(force), native.stream_of(f_0),
            )
        )  # fmt: skip
        return force[: vs.d].to(self.compute_dtype)

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f_0, f_1, bc_mask, missing_mask):
        return Field.wrap(self._run(f_0, f_1, bc_mask, missing_mask))

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_0, f_1, bc_mask, missing_mask):
        return self._run(f_0, f_1, bc_mask, missing_mask).cpu().numpy()  # reference returns force.numpy()[0]
