// Explicit instantiation of the fused step kernel family for the register-lean KBC formulation (the KBC default; D3Q27, D2Q9 and the
// 2-D slab axis order).  This unit alone is compiled with -fmad=true: the lean form is the FAST form (reciprocal-based divisions,
// explicit fused operations, held to the 1e-5 tolerance), and implicit contraction of its remaining multiply-add pairs is worth 6 % of
// its throughput on B200; every other unit — the BGK chain and the literal KBC parity form among them — rounds after every operation.
#include "step_kernel.cuh"

namespace xlbn {
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_KBC | kLeanKbc)
XLBN_DEFINE_STEP_DISPATCH(D2Q9, XLBN_KBC | kLeanKbc)
XLBN_DEFINE_STEP_DISPATCH(D2Q9X, XLBN_KBC | kLeanKbc)
// ... and under ForcedCollision (ExactDifference added in the last pass)
XLBN_DEFINE_STEP_DISPATCH(D3Q27, XLBN_KBC | kLeanKbc | XLBN_COLLISION_FORCED)
XLBN_DEFINE_STEP_DISPATCH(D2Q9, XLBN_KBC | kLeanKbc | XLBN_COLLISION_FORCED)
}  // namespace xlbn
