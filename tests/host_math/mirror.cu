// TEST HARNESS — compiles the library's per-cell algebra (xlb_b200/csrc/lbm_math.cuh: the very functions the CUDA
// kernels inline) for the HOST, so that tests can run them on the CPU against the oracle when no GPU is available.
// Built by tests/test_host_mirror_math.py with   nvcc -DXLBN_HOST_MIRROR -shared ...   into tests/host_math/_build/.
// It checks arithmetic and control flow of the cell functions; memory indexing of the kernels is not exercised here.
// Arrays are cell-major: f[cell][q].
#define XLBN_HOST_MIRROR 1
#include "../../xlb_b200/csrc/lbm_math.cuh"

using namespace xlbn;

namespace {

template <class L, int COLL, class TC, bool FAST>
void collide_all(long n, const TC* f, TC* out, double omega, const double* force, double smagorinsky) {
  for (long i = 0; i < n; ++i) {
    TC c[L::Q];
    for (int l = 0; l < L::Q; ++l) c[l] = f[i * L::Q + l];
    if constexpr (COLL <= XLBN_KBC) collide_cell<L, COLL, TC, FAST>(c, (TC)omega);
    else collide_cell_ext<L, COLL, TC, FAST>(c, (TC)omega, force, smagorinsky);
    for (int l = 0; l < L::Q; ++l) out[i * L::Q + l] = c[l];
  }
}

template <class L, int COLL>
int collide_typed(int compute, int fast, long n, const void* f, void* out, double omega, const double* force, double smagorinsky) {
  if (compute == XLBN_F32) {
    if (fast) collide_all<L, COLL, float, true>(n, (const float*)f, (float*)out, omega, force, smagorinsky);
    else collide_all<L, COLL, float, false>(n, (const float*)f, (float*)out, omega, force, smagorinsky);
  } else {
    collide_all<L, COLL, double, false>(n, (const double*)f, (double*)out, omega, force, smagorinsky);
  }
  return 0;
}

template <class L, class TC>
void zouhe_all(int kind, long n, const TC* f, const TC* aux, const uint32_t* miss, TC* out) {
  for (long i = 0; i < n; ++i) {
    TC c[L::Q];
    for (int l = 0; l < L::Q; ++l) c[l] = f[i * L::Q + l];
    bc_zouhe<L, TC>(kind, aux[i], miss[i], c);
    for (int l = 0; l < L::Q; ++l) out[i * L::Q + l] = c[l];
  }
}

}  // namespace

extern "C" {

// collision = xlbn_collision code (base | XLBN_COLLISION_FORCED); returns -1 for combinations the library does not build
int mirror_collide(int lattice, int collision, int compute, int fast, long n, const void* f, void* out, double omega, const double* force,
                   double smagorinsky) {
  constexpr int F = XLBN_COLLISION_FORCED;
#define CASE(LAT, TAG, COLL) \
  if (lattice == TAG && collision == (COLL)) return collide_typed<LAT, (COLL)>(compute, fast, n, f, out, omega, force, smagorinsky);
  CASE(D3Q19, XLBN_D3Q19, XLBN_BGK)
  CASE(D3Q19, XLBN_D3Q19, XLBN_BGK | F)
  CASE(D3Q19, XLBN_D3Q19, XLBN_SMAGORINSKY_LES_BGK)
  CASE(D3Q19, XLBN_D3Q19, XLBN_SMAGORINSKY_LES_BGK | F)
  CASE(D3Q27, XLBN_D3Q27, XLBN_BGK)
  CASE(D3Q27, XLBN_D3Q27, XLBN_KBC)
  CASE(D3Q27, XLBN_D3Q27, XLBN_BGK | F)
  CASE(D3Q27, XLBN_D3Q27, XLBN_KBC | F)
  CASE(D3Q27, XLBN_D3Q27, XLBN_SMAGORINSKY_LES_BGK)
  CASE(D3Q27, XLBN_D3Q27, XLBN_SMAGORINSKY_LES_BGK | F)
  CASE(D2Q9, XLBN_D2Q9, XLBN_BGK)
  CASE(D2Q9, XLBN_D2Q9, XLBN_KBC)
  CASE(D2Q9, XLBN_D2Q9, XLBN_BGK | F)
  CASE(D2Q9, XLBN_D2Q9, XLBN_KBC | F)
#undef CASE
  return -1;
}

// Zou-He / Regularized functional on post-stream populations; miss = missing-direction bitmask per cell
int mirror_bc_zouhe(int lattice, int compute, int kind, long n, const void* f, const void* aux, const uint32_t* miss, void* out) {
#define CASE(LAT, TAG)                                                                                    \
  if (lattice == TAG) {                                                                                   \
    if (compute == XLBN_F32) zouhe_all<LAT, float>(kind, n, (const float*)f, (const float*)aux, miss, (float*)out);   \
    else zouhe_all<LAT, double>(kind, n, (const double*)f, (const double*)aux, miss, (double*)out);                    \
    return 0;                                                                                             \
  }
  CASE(D3Q19, XLBN_D3Q19)
  CASE(D3Q27, XLBN_D3Q27)
  CASE(D2Q9, XLBN_D2Q9)
#undef CASE
  return -1;
}

}  // extern "C"
