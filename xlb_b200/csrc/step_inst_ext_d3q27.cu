// Explicit instantiation of the fused step kernel for D3Q27 with the extended collision models:
// SmagorinskyLESBGK and the ForcedCollision / ExactDifference wrapper.  One cell per thread, all precision policies.
#include "step_kernel.cuh"

namespace xlbn {
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_BGK | XLBN_COLLISION_FORCED)
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_KBC | XLBN_COLLISION_FORCED)
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_SMAGORINSKY_LES_BGK)
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_SMAGORINSKY_LES_BGK | XLBN_COLLISION_FORCED)
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_KBC | kExactKbc)  // parity form: the reference's roundings (cells_per_thread = 300)
}  // namespace xlbn
