// Boundary masks: bc_mask (uint8 id per cell) and missing_mask (bool per direction per cell) from index lists.
// Replaces IndicesBoundaryMasker (reference: xlb/operator/boundary_masker/indices_boundary_masker.py); both of the
// reference's algorithms are reproduced bit-exactly (they differ only on BC-free domain-face cells):
//   XLBN_MASK_WARP  L103-224: per index, bc_mask[idx] = id; missing[l, idx] if idx - c_l leaves the domain; for
//                   strictly-interior indices of needs_padding BCs: missing[l, idx + c_l] and bc_mask[idx + c_l] = id.
//   XLBN_MASK_JAX   L45-101 : ids written per BC; BCs with needs_padding and ANY interior index mark their cells solid
//                   and push their id to all neighbours; finally missing[l, x] = outside_or_solid(x - c_l) for EVERY cell.
// Slab-aware: indices are GLOBAL; each rank keeps what lands in its slab (one solid-halo cell deep for the JAX mode).
#include "lbm_math.cuh"

namespace xlbn {

struct MaskGeom {
  int g[3];      // global extents
  int s[3];      // global coordinate of local cell (0,0,0)
  int n[3];      // local extents
  long long cells;
};

__device__ __forceinline__ bool in_global(const MaskGeom& m, int x, int y, int z) {
  return x >= 0 && x < m.g[0] && y >= 0 && y < m.g[1] && z >= 0 && z < m.g[2];
}
__device__ __forceinline__ long long local_index(const MaskGeom& m, int x, int y, int z) {  // -1 if outside the slab
  x -= m.s[0];
  y -= m.s[1];
  z -= m.s[2];
  if (x < 0 || x >= m.n[0] || y < 0 || y >= m.n[1] || z < 0 || z >= m.n[2]) return -1;
  return ((long long)x * m.n[1] + y) * m.n[2] + z;
}
// index into the solid scratch: local extents + 1 halo cell on each side of EVERY axis
__device__ __forceinline__ long long solid_index(const MaskGeom& m, int x, int y, int z) {
  x -= m.s[0] - 1;
  y -= m.s[1] - 1;
  z -= m.s[2] - 1;
  if (x < 0 || x >= m.n[0] + 2 || y < 0 || y >= m.n[1] + 2 || z < 0 || z >= m.n[2] + 2) return -1;
  return ((long long)x * (m.n[1] + 2) + y) * (m.n[2] + 2) + z;
}

template <class L>
__device__ __forceinline__ int cphys(int a, int l) {
  return a < L::D ? L::c(a < L::D ? a : 0, l) : 0;
}

template <class L>
__global__ void mask_indices_kernel(int mode, const int32_t* idx, long long n_idx, int bc_id, int needs_padding, MaskGeom m, uint8_t* bc_mask,
                                    uint8_t* missing, uint8_t* solid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_idx) return;
  const int x = idx[i], y = idx[n_idx + i], z = idx[2 * n_idx + i];
  if (!in_global(m, x, y, z)) return;  // indices_boundary_masker.py:131 (check_index_bounds)
  const long long here = local_index(m, x, y, z);
  if (here >= 0) bc_mask[here] = (uint8_t)bc_id;
  const bool interior = x > 0 && x < m.g[0] - 1 && y > 0 && y < m.g[1] - 1 && (L::D == 2 || (z > 0 && z < m.g[2] - 1));
  if (mode == XLBN_MASK_WARP) {
    XLBN_FOR(L::Q, l)
      const int cx = cphys<L>(0, l), cy = cphys<L>(1, l), cz = cphys<L>(2, l);
      if (!in_global(m, x - cx, y - cy, z - cz)) {
        if (here >= 0) missing[(long long)l * m.cells + here] = 1;
      } else if (needs_padding && interior) {
        const long long push = local_index(m, x + cx, y + cy, z + cz);
        if (push >= 0) {
          missing[(long long)l * m.cells + push] = 1;
          bc_mask[push] = (uint8_t)bc_id;
        }
      }
    XLBN_END
  } else if (needs_padding) {  // JAX mode; the flag already includes "any index of this BC is interior"
    const long long sh = solid_index(m, x, y, z);
    if (sh >= 0) solid[sh] = 1;
    XLBN_FOR(L::Q, l)
      const int px = x + cphys<L>(0, l), py = y + cphys<L>(1, l), pz = z + cphys<L>(2, l);
      if (in_global(m, px, py, pz)) {
        const long long push = local_index(m, px, py, pz);
        if (push >= 0) bc_mask[push] = (uint8_t)bc_id;
      }
    XLBN_END
  }
}

template <class L>
__global__ void mask_finalize_jax_kernel(MaskGeom m, uint8_t* missing, const uint8_t* solid, const uint8_t* incoming) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.cells) return;
  const int z = (int)(i % m.n[2]) + m.s[2];
  const int y = (int)((i / m.n[2]) % m.n[1]) + m.s[1];
  const int x = (int)(i / ((long long)m.n[2] * m.n[1])) + m.s[0];
  XLBN_FOR(L::Q, l)
    const int sx = x - cphys<L>(0, l), sy = y - cphys<L>(1, l), sz = z - cphys<L>(2, l);
    bool miss = !in_global(m, sx, sy, sz);
    if (!miss) {
      const long long sh = solid_index(m, sx, sy, sz);
      miss = sh >= 0 && solid[sh] != 0;
      // entries the caller's mask already held travel with the stream as well (the reference pads the INCOMING mask and streams it,
      // indices_boundary_masker.py:56-63, 92): only sources inside this slab can be seen here
      if (!miss && incoming) {
        const int lx = sx - m.s[0], ly = sy - m.s[1], lz = sz - m.s[2];
        if (lx >= 0 && lx < m.n[0] && ly >= 0 && ly < m.n[1] && lz >= 0 && lz < m.n[2])
          miss = incoming[(long long)l * m.cells + ((long long)lx * m.n[1] + ly) * m.n[2] + lz] != 0;
      }
    }
    missing[(long long)l * m.cells + i] = miss ? 1 : 0;
  XLBN_END
}

__global__ void pack_missing_kernel(int q, const uint8_t* missing, uint32_t* bits, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t b = 0;
  for (int l = 0; l < q; ++l) b |= (missing[(long long)l * n + i] ? 1u : 0u) << l;
  bits[i] = b;
}

static int make_geom(const int32_t g[3], const int32_t s[3], const int32_t n[3], MaskGeom* m) {
  if (!g || !s || !n) return fail(XLBN_E_ARG, "mask: NULL geometry");
  for (int a = 0; a < 3; ++a) {
    if (g[a] <= 0 || n[a] <= 0 || s[a] < 0 || s[a] + n[a] > g[a])
      return fail(XLBN_E_SHAPE, "mask: axis %d: global %d start %d local %d", a, g[a], s[a], n[a]);
    m->g[a] = g[a];
    m->s[a] = s[a];
    m->n[a] = n[a];
  }
  m->cells = (long long)n[0] * n[1] * n[2];
  return 0;
}

}  // namespace xlbn

using namespace xlbn;

extern "C" {

int xlbn_mask_indices(int lattice, int mode, const int32_t* indices, long long n, int bc_id, int needs_padding, const int32_t global_dims[3],
                      const int32_t start[3], const int32_t local_dims[3], uint8_t* bc_mask, uint8_t* missing, uint8_t* solid, void* stream) {
  XLBN_RANGE("xlbn_mask_indices");
  MaskGeom m;
  if (int e = make_geom(global_dims, start, local_dims, &m)) return e;
  if (n < 0) return fail(XLBN_E_ARG, "mask: negative index count");
  if (n == 0) return 0;
  if (!indices || !bc_mask || !missing) return fail(XLBN_E_ARG, "mask: NULL array");
  if (bc_id < 0 || bc_id > 255) return fail(XLBN_E_ARG, "mask: bc id %d outside uint8", bc_id);
  if (mode != XLBN_MASK_WARP && mode != XLBN_MASK_JAX) return fail(XLBN_E_ARG, "mask: unknown mode %d", mode);
  if (mode == XLBN_MASK_JAX && needs_padding && !solid) return fail(XLBN_E_ARG, "mask: JAX mode needs the solid scratch array");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (lattice) {
    case XLBN_D2Q9: mask_indices_kernel<D2Q9><<<blocks, 256, 0, st>>>(mode, indices, n, bc_id, needs_padding, m, bc_mask, missing, solid); break;
    case XLBN_D3Q19: mask_indices_kernel<D3Q19><<<blocks, 256, 0, st>>>(mode, indices, n, bc_id, needs_padding, m, bc_mask, missing, solid); break;
    case XLBN_D3Q27: mask_indices_kernel<D3Q27><<<blocks, 256, 0, st>>>(mode, indices, n, bc_id, needs_padding, m, bc_mask, missing, solid); break;
    default: return fail(XLBN_E_ARG, "unknown lattice %d", lattice);
  }
  XLBN_LAUNCH_OK("mask_indices_kernel");
  return 0;
}

int xlbn_mask_finalize_jax(int lattice, const int32_t global_dims[3], const int32_t start[3], const int32_t local_dims[3], uint8_t* missing,
                           const uint8_t* solid, const uint8_t* incoming, void* stream) {
  XLBN_RANGE("xlbn_mask_finalize_jax");
  MaskGeom m;
  if (int e = make_geom(global_dims, start, local_dims, &m)) return e;
  if (!missing || !solid) return fail(XLBN_E_ARG, "mask finalize: NULL array");
  if (incoming == missing) return fail(XLBN_E_ARG, "mask finalize: `incoming` must be a copy of the caller's mask, not the output array");
  const unsigned blocks = (unsigned)((m.cells + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  switch (lattice) {
    case XLBN_D2Q9: mask_finalize_jax_kernel<D2Q9><<<blocks, 256, 0, st>>>(m, missing, solid, incoming); break;
    case XLBN_D3Q19: mask_finalize_jax_kernel<D3Q19><<<blocks, 256, 0, st>>>(m, missing, solid, incoming); break;
    case XLBN_D3Q27: mask_finalize_jax_kernel<D3Q27><<<blocks, 256, 0, st>>>(m, missing, solid, incoming); break;
    default: return fail(XLBN_E_ARG, "unknown lattice %d", lattice);
  }
  XLBN_LAUNCH_OK("mask_finalize_jax_kernel");
  return 0;
}

int xlbn_pack_missing(int q, const uint8_t* missing, uint32_t* bits, long long n_cells, void* stream) {
  XLBN_RANGE("xlbn_pack_missing");
  if (q <= 0 || q > 32) return fail(XLBN_E_ARG, "pack_missing: q = %d", q);
  if (!missing || !bits || n_cells < 0) return fail(XLBN_E_ARG, "pack_missing: bad argument");
  if (n_cells == 0) return 0;
  pack_missing_kernel<<<(unsigned)((n_cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q, missing, bits, n_cells);
  XLBN_LAUNCH_OK("pack_missing_kernel");
  return 0;
}

}  // extern "C"
