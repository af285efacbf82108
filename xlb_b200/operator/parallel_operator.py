"""Kept for import compatibility with reference xlb/operator/parallel_operator.py (a duplicate of
xlb/distribute/distribute.py built on jax.pmap-era APIs).  The x-slab decomposition of this framework lives in
xlb_b200/distribute; this wrapper simply forwards to it."""


class ParallelOperator:
    def __init__(self, grid, func, velocity_set, num_results=1, ops="permute"):
        from xlb_b200.distribute import distribute

        self.grid, self.func, self.velocity_set = grid, func, velocity_set
        self._wrapped = distribute(func, grid, velocity_set, num_results=num_results, ops=ops)

    def __call__(self, *args, **kwargs):
        return self._wrapped(*args, **kwargs)
