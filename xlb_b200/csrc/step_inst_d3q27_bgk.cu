// Explicit instantiation of the fused step kernel family for D3Q27 / XLBN_BGK (all precision policies, all V).
#include "step_kernel.cuh"

namespace xlbn {
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_BGK)
}  // namespace xlbn
