// Explicit instantiation of the fused step kernel family for D2Q9 in the x-slab axis order (D2Q9X, lattice.cuh): BGK and KBC, all
// precision policies.  Used by xlbn_step when a 2-D call carries a halo handle or a partial x range.
#include "step_kernel.cuh"

namespace xlbn {
XLBN_DEFINE_STEP_DISPATCH(D2Q9X, XLBN_BGK)
XLBN_DEFINE_STEP_DISPATCH(D2Q9X, XLBN_KBC)
}  // namespace xlbn
