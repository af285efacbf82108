// Explicit instantiation of the fused step kernel family for D2Q9 / XLBN_KBC (all precision policies, all V).
#include "step_kernel.cuh"

namespace xlbn {
XLBN_DEFINE_STEP_DISPATCH(D2Q9, XLBN_KBC)
}  // namespace xlbn
