"""x-slab distribution of operators over the GPUs of one box.

Reference: xlb/distribute/distribute.py:10-105.  There, `distribute(stepper, grid, velocity_set)` wraps the stepper's
`stream` (or the whole stepper) in a `shard_map` whose body streams with a LOCAL periodic roll and then repairs the two
wrongly wrapped face planes with two `lax.ppermute` ring collectives (L23-44).

Here the decomposition is a property of the grid (one process per GPU under torch.distributed, xlb_b200/grid/grid.py):
 * a stepper on a slab grid already exchanges its halo inside the fused step (PeerHalo), so `distribute(stepper, ...)`
   returns the stepper unchanged — kept so that reference scripts (examples/performance/mlups_3d.py:66-71) run as is;
 * any other operator whose result is a population field (e.g. `Stream`) gets the reference's post-fix-up semantics:
   run locally, then swap `result[right_indices, :1]` / `result[left_indices, -1:]` with the ring neighbours
   (point-to-point over NCCL on GPUs, gloo in the CPU tests of the host logic).
With a single slab both return the operator itself.
"""


from xlb_b200.distribute.halo import exchange_wrapped_faces


def distribute_operator(operator, grid, velocity_set, num_results=1, ops="permute"):
    if ops != "permute":
        raise NotImplementedError(f"Operation {ops} not implemented")
    if grid.nDevices == 1:
        return operator

    def _sharded_operator(*args):
        result = operator(*args)
        if num_results == 1:
            return exchange_wrapped_faces(result, velocity_set, grid.rank, grid.nDevices)
        return tuple(exchange_wrapped_faces(r, velocity_set, grid.rank, grid.nDevices) for r in result)

    return _sharded_operator


def distribute(operator, grid, velocity_set, num_results=1, ops="permute"):
    from xlb_b200.operator.stepper import IncompressibleNavierStokesStepper

    if isinstance(operator, IncompressibleNavierStokesStepper):
        if operator.grid is not grid:
            operator.grid = grid
        return operator
    return distribute_operator(operator, grid, velocity_set, num_results=num_results, ops=ops)
