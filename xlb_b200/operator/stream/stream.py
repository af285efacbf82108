"""Pull streaming with periodic wrap: f_out[l, x] = f_in[l, x - c_l].

Reference: xlb/operator/stream/stream.py — JAX ``(f) -> f`` L18-51, Warp ``(f_0, f_1) -> f_1`` L101-114.
Native: xlbn_stream (xlb_b200/csrc/ops.cu).  Works on any element type incl. bool masks (the JAX masker streams a
bool array, indices_boundary_masker.py:94).
"""

import torch

from xlb_b200 import native
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.field import Field
from xlb_b200.operator.operator import Operator
from xlb_b200.operator._util import to_device_field


class Stream(Operator):
    def _run(self, f_in, f_out):
        native.require_cuda(f_in, "f")
        native.require_cuda(f_out, "f_out")
        if f_in.shape != f_out.shape or f_in.dtype != f_out.dtype:
            raise ValueError("Stream: input and output must have the same shape and dtype")
        if f_in.shape[0] != self.velocity_set.q:
            raise ValueError(f"Stream: leading axis must be q = {self.velocity_set.q}")
        dims = native.dims_of(f_in, self.velocity_set.d)
        native.check(
            native.lib().xlbn_stream(self._lattice, native.ptr(f_in), native.ptr(f_out), native.dtype_code(f_in.dtype), native.int3(dims), native.stream_of(f_in))
        )
        return f_out

    @Operator.register_backend(ComputeBackend.JAX)
    def jax_implementation(self, f):
        f = to_device_field(f)
        return self._run(f, Field.wrap(torch.empty_like(f)))

    @Operator.register_backend(ComputeBackend.WARP)
    def warp_implementation(self, f_0, f_1):
        return self._run(f_0, f_1)
