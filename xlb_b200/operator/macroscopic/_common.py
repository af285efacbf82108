"""Shared native call of the moment operators (xlbn_macroscopic / xlbn_first_moment / xlbn_second_moment)."""

from xlb_b200 import native


def check_f(op, f):
    native.require_cuda(f, "f")
    if f.shape[0] != op.velocity_set.q:
        raise ValueError(f"{type(op).__name__}: leading axis of f must be q = {op.velocity_set.q}, got {f.shape[0]}")
    return native.dims_of(f, op.velocity_set.d)


def run_macroscopic(op, f, rho, u):
    dims = check_f(op, f)
    for name, t, card in (("rho", rho, 1), ("u", u, op.velocity_set.d)):
        if t is not None:
            native.require_cuda(t, name)
            if t.shape[0] != card or t.shape[1:] != f.shape[1:]:
                raise ValueError(f"{type(op).__name__}: {name} has shape {tuple(t.shape)}")
    native.check(
        native.lib().xlbn_macroscopic(
            op._lattice, op._compute_code, native.ptr(f), native.dtype_code(f.dtype),
            native.ptr(rho), native.dtype_code(rho.dtype) if rho is not None else 0,
            native.ptr(u), native.dtype_code(u.dtype) if u is not None else 0,
            native.int3(dims), native.stream_of(f),
        )
    )  # fmt: skip
