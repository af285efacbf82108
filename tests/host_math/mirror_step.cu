// TEST HARNESS — the fused step's SOURCE (xlb_b200/csrc/step_kernel.cuh: fill_step_params, step_body, bc_tail, bc_cell,
// store_cells and the cell algebra they inline) compiled for the HOST and driven by plain loops instead of a CUDA launch.
// The loops below do what step_kernel's prologue does (thread coordinates -> x, y, z0; x-plane class); everything else —
// pointer tables, pull addressing incl. periodic wrap and ghost planes, boundary dispatch, collision, stores incl. the
// peer-plane stores — is the shipped code, incl. the packed-pair path (cells_per_thread 102) and the half2-state path (202),
// whose packed instructions and warp votes have one-lane host twins (common.cuh).  Not covered: launch geometry and anything
// that only exists on a GPU (coalescing, scheduling, races).
// Built by tests/test_host_mirror_step.py: one object per (lattice, collision) — -DMIRROR_LAT=D3Q27 -DMIRROR_TAG=XLBN_D3Q27
// -DMIRROR_COLL=5 -DMIRROR_NAME=mirror_step_2_5, compiled in parallel — plus error.cu, linked into
// tests/host_math/_build/libmirror_step.so.
#define XLBN_HOST_MIRROR 1
#if !defined(MIRROR_LAT) || !defined(MIRROR_TAG) || !defined(MIRROR_COLL) || !defined(MIRROR_NAME)
#error "compile with -DMIRROR_LAT=<lattice type> -DMIRROR_TAG=<xlbn_lattice> -DMIRROR_COLL=<collision code> -DMIRROR_NAME=<symbol>"
#endif
#include "../../xlb_b200/csrc/step_kernel.cuh"

using namespace xlbn;

namespace {

template <class L, int COLL, class TC, class TS, int V>
int host_step(const StepCall& c) {
  StepParams<TS> p;
  if (int e = fill_step_params<L, TS>(c, p)) return e;
  if (c.nz % V) return fail(XLBN_E_SHAPE, "mirror: nz %% V != 0");
  for (int x = c.x_begin; x < c.x_begin + c.x_count; ++x) {
    const bool first = (x == 0), last = (x == p.nx - 1);  // as step_kernel
    for (int y = 0; y < p.ny; ++y)
      for (int z0 = 0; z0 < p.nz; z0 += V) {
        if (!first && !last) step_body<L, COLL, TC, TS, V, 0>(p, x, y, z0);
        else if (first && !last) step_body<L, COLL, TC, TS, V, 1>(p, x, y, z0);
        else if (last && !first) step_body<L, COLL, TC, TS, V, 2>(p, x, y, z0);
        else step_body<L, COLL, TC, TS, V, 3>(p, x, y, z0);
      }
  }
  return 0;
}

// MODE 1 (packed fp32x2 pair path, two cells per thread) and MODE 2 (half2-state pair path): the other bodies of step_kernel
template <class L, int COLL, class TS, int MODE>
int host_step_pair(const StepCall& c) {
  StepParams<TS> p;
  if (int e = fill_step_params<L, TS>(c, p)) return e;
  if (c.nz % 2) return fail(XLBN_E_SHAPE, "mirror: pair paths need an even nz");
  if constexpr (MODE != 1) {  // what bc_precompute_kernel does on the device: EquilibriumBC cells' constant update
    BcEntry* table = c.table_rw;
    for (int id = 0; id < 256; ++id) {
      if (table[id].kind != XLBN_BC_EQUILIBRIUM) continue;
      float u[L::D], f[L::Q];
      XLBN_FOR(L::D, d) u[d] = (float)table[id].u[d]; XLBN_END
      equilibrium<L, float>((float)table[id].rho, u, f);
      collide_cell<L, COLL, float, kFast<COLL, float>>(f, (float)c.omega);
      XLBN_FOR(L::Q, l) table[id].eq_out[l] = f[l]; XLBN_END
    }
  }
  for (int x = c.x_begin; x < c.x_begin + c.x_count; ++x) {
    const bool first = (x == 0), last = (x == p.nx - 1);
    for (int y = 0; y < p.ny; ++y)
      for (int z0 = 0; z0 < p.nz; z0 += 2) {
        if constexpr (MODE == 1) {
          if (!first && !last) step_body_pk<L, COLL, TS, 2, 0>(p, x, y, z0);
          else if (first && !last) step_body_pk<L, COLL, TS, 2, 1>(p, x, y, z0);
          else if (last && !first) step_body_pk<L, COLL, TS, 2, 2>(p, x, y, z0);
          else step_body_pk<L, COLL, TS, 2, 3>(p, x, y, z0);
        } else if constexpr (MODE == 2) {
          if (!first && !last) step_body_h2<L, 0>(p, x, y, z0);
          else if (first && !last) step_body_h2<L, 1>(p, x, y, z0);
          else if (last && !first) step_body_h2<L, 2>(p, x, y, z0);
          else step_body_h2<L, 3>(p, x, y, z0);
        } else {  // MODE 5: boundary warps split by kind
          if (!first && !last) step_body_h2<L, 0, true>(p, x, y, z0);
          else if (first && !last) step_body_h2<L, 1, true>(p, x, y, z0);
          else if (last && !first) step_body_h2<L, 2, true>(p, x, y, z0);
          else step_body_h2<L, 3, true>(p, x, y, z0);
        }
      }
  }
  return 0;
}

// The tile kernel (csrc/step_tile.cuh; cells_per_thread 402) with its asynchronous machinery replaced by what it amounts to: the
// producer's copy plan executed with memcpy into an input stage, the 256 consumer threads of the tile run one after the other
// (tile_load + tile_compute: the shipped per-thread code, which stores to the destination itself).
template <class L, int CELLS>
int host_step_tile(const StepCall& c) {
  constexpr int kTileCells = CELLS, kTileConsumers = TileDims<CELLS>::kConsumers, kTileRowBytes = TileDims<CELLS>::kRowBytes;
  StepParams<__half> p;
  if (int e = fill_step_params<L, __half>(c, p)) return e;
  if (!tile_eligible<L, CELLS>(p, c.f0, c.f1, c.ghost_lo, c.ghost_hi, c.out_lo != nullptr || c.out_hi != nullptr)) return fail(XLBN_E_SHAPE, "mirror: shape not eligible for the tile kernel");
  BcEntry* table = c.table_rw;  // bc_precompute_kernel
  for (int id = 0; id < 256; ++id) {
    if (table[id].kind != XLBN_BC_EQUILIBRIUM) continue;
    float u[L::D], f[L::Q];
    XLBN_FOR(L::D, d) u[d] = (float)table[id].u[d]; XLBN_END
    equilibrium<L, float>((float)table[id].rho, u, f);
    collide_cell<L, XLBN_BGK, float, false>(f, (float)c.omega);
    XLBN_FOR(L::Q, l) table[id].eq_out[l] = f[l]; XLBN_END
  }
  using C = TileCfg<L, CELLS, 1>;
  const int rows = kTileCells / p.nz, tiles_per_plane = p.ny / rows, n_tiles = tiles_per_plane * c.x_count;
  static unsigned char in[C::kInBytes];
  static TileEqTable eq;
  tile_eq_table_fill<L>(p, eq);  // the kernel's prologue (one thread per CTA)
  for (int tile = 0; tile < n_tiles; ++tile) {
    const TileGeom g = tile_geom(p, tile, rows, tiles_per_plane);
    memset(in, 0xff, sizeof(in));
    for (int l = 0; l < L::Q; ++l) {  // the producer warp, lane l
      const TileRuns<__half> r = tile_runs<L, __half>(p, l, L::ck(0, l), L::ck(1, l), g, rows);
      if (r.n0 + r.n1 != (unsigned)kTileCells) return fail(XLBN_E_SHAPE, "mirror: copy plan covers %u of %d elements", r.n0 + r.n1, kTileCells);
      if ((reinterpret_cast<uintptr_t>(r.src0) % 16) || ((r.n0 * 2u) % 16) || (r.n1 && (reinterpret_cast<uintptr_t>(r.src1) % 16))) return fail(XLBN_E_SHAPE, "mirror: bulk copy not 16-byte aligned");
      memcpy(in + l * kTileRowBytes, r.src0, r.n0 * 2u);
      if (r.n1) memcpy(in + l * kTileRowBytes + r.n0 * 2u, r.src1, r.n1 * 2u);
    }
    memcpy(in + L::Q * kTileRowBytes, p.bc + g.cell0, kTileCells);
    for (unsigned t = 0; t < (unsigned)kTileConsumers; ++t) {  // the consumer threads
      __half2 h[L::Q];
      unsigned ids;
      tile_load<L, CELLS>(p, reinterpret_cast<const uint32_t*>(in), in + L::Q * kTileRowBytes, t, h, ids);
      tile_compute<L, CELLS>(p, eq, h, ids, t, g);
    }
  }
  return 0;
}

// The scalar tile kernel (cells_per_thread 501): copy plan as scalars (tile_runs) executed with memcpy, then the 512 consumer threads of
// the tile one after the other (tile1_load + tile1_finish: the shipped per-thread code).
template <class L, int COLL, class TC, class TS>
int host_step_tile1(const StepCall& c) {
  StepParams<TS> p;
  if (int e = fill_step_params<L, TS>(c, p)) return e;
  if (!tile1_eligible<L, TS>(p, c.f0, c.f1, c.ghost_lo, c.ghost_hi, c.out_lo != nullptr || c.out_hi != nullptr)) return fail(XLBN_E_SHAPE, "mirror: shape not eligible for the scalar tile kernel");
  const int rows = kT1Cells / p.nz, tiles_per_plane = p.ny / rows, n_tiles = tiles_per_plane * c.x_count;
  static TS in[L::Q * kT1Cells];
  static unsigned char id_row[kT1Cells];
  const unsigned nz = (unsigned)p.nz;
  for (int tile = 0; tile < n_tiles; ++tile) {
    const TileGeom g = tile_geom(p, tile, rows, tiles_per_plane);
    memset(in, 0xff, sizeof(in));
    for (int l = 0; l < L::Q; ++l) {  // the producer warp, lane l
      const TileRuns<TS> r = tile_runs<L, TS>(p, l, L::ck(0, l), L::ck(1, l), g, rows);
      if (r.n0 + r.n1 != (unsigned)kT1Cells) return fail(XLBN_E_SHAPE, "mirror: copy plan covers %u of 512 elements", r.n0 + r.n1);
      if ((reinterpret_cast<uintptr_t>(r.src0) % 16) || ((r.n0 * sizeof(TS)) % 16) || (r.n1 && (reinterpret_cast<uintptr_t>(r.src1) % 16))) return fail(XLBN_E_SHAPE, "mirror: bulk copy not 16-byte aligned");
      memcpy(in + l * kT1Cells, r.src0, r.n0 * sizeof(TS));
      if (r.n1) memcpy(in + l * kT1Cells + r.n0, r.src1, r.n1 * sizeof(TS));
    }
    memcpy(id_row, p.bc + g.cell0, kT1Cells);
    for (unsigned t = 0; t < (unsigned)kT1Cells; ++t) {  // the consumer threads
      const unsigned z = t & (nz - 1u);
      const unsigned tm = (z == 0u) ? t + nz - 1u : t - 1u, tp = (z + 1u == nz) ? t + 1u - nz : t + 1u;
      TC f[1][L::Q];
      Pack<uint8_t, 1> ids;
      tile1_load<L, TC, TS>(in, id_row, t, tm, tp, f, ids);
      tile1_finish<L, COLL, TC, TS>(p, g, t, z, (TC)p.omega, f, ids);
    }
  }
  return 0;
}

template <class L, int COLL, class TC, class TS>
int host_step_v(const StepCall& c) {
  if constexpr ((COLL == XLBN_BGK || (kExtCollision<COLL> && (COLL & (kLeanKbc | kExactKbc)) == 0)) && sizeof(TS) == 4 && L::D == 3)
    if (c.requested_v == 501) return host_step_tile1<L, COLL, TC, TS>(c);
  if constexpr (COLL == XLBN_BGK && sizeof(TC) == 4 && sizeof(TS) == 2 && L::D == 3)
    if (c.requested_v == 404) return host_step_tile<L, 512>(c);
  if constexpr (COLL == XLBN_BGK && sizeof(TC) == 4 && sizeof(TS) == 2 && L::D == 3)
    if (c.requested_v == 402) return host_step_tile<L, 1024>(c);  // 1024-cell tiles (the library falls back to 512 when they do not fit)
  if (c.requested_v == 1) return host_step<L, COLL, TC, TS, 1>(c);
  if constexpr (!kExtCollision<COLL> && sizeof(TC) == 4) {
    if (c.requested_v == 102) return host_step_pair<L, COLL, TS, 1>(c);
    if constexpr (sizeof(TS) == 2 && COLL == XLBN_BGK)
      if (c.requested_v == 202 || c.requested_v == 203) return c.requested_v == 202 ? host_step_pair<L, COLL, TS, 2>(c) : host_step_pair<L, COLL, TS, 5>(c);
  }
  if constexpr (!kExtCollision<COLL>) {  // the library builds the extended operators with one cell per thread only
    if (c.requested_v == 2) return host_step<L, COLL, TC, TS, 2>(c);
    if constexpr (sizeof(TS) <= 4)
      if (c.requested_v == 4) return host_step<L, COLL, TC, TS, 4>(c);
  }
  return fail(XLBN_E_ARG, "mirror: cells per thread %d", c.requested_v);
}

template <class L, int COLL>
int host_step_policy(const StepCall& c) {
  if (c.compute_dtype == XLBN_F32 && c.store_dtype == XLBN_F32) return host_step_v<L, COLL, float, float>(c);
  if (c.compute_dtype == XLBN_F32 && c.store_dtype == XLBN_F16) return host_step_v<L, COLL, float, __half>(c);
  if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F64) return host_step_v<L, COLL, double, double>(c);
  if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F32) return host_step_v<L, COLL, double, float>(c);
  if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F16) return host_step_v<L, COLL, double, __half>(c);
  return fail(XLBN_E_DTYPE, "mirror: policy");
}

}  // namespace

extern "C" {

#ifdef MIRROR_DEFINE_ERROR
const char* mirror_last_error(void) { return error_buffer(); }
#endif

// One step of the fused kernel's source on the host.  Arguments as xlbn_step + xlbn_stepper_desc, flattened; all pointers are
// HOST pointers.  dims are the KERNEL extents (2-D: (1, nx, ny), as xlbn_step passes them).  bc_kind / bc_rho / bc_u: 256-entry
// tables indexed by bc id.  ghost_* / out_*: x-slab ghost planes as in xlbn_step's halo (NULL: periodic in x).
int MIRROR_NAME(int lattice, int collision, int compute_dtype, int store_dtype, int cells_per_thread, const void* f0, void* f1, const uint8_t* bc_mask,
                const uint32_t* missing_bits, const int32_t* bc_kind, const double* bc_rho, const double* bc_u, const int32_t dims[3], int x_begin,
                int x_count, double omega, const double* force, double smagorinsky, const void* ghost_lo, const void* ghost_hi, void* out_lo,
                void* out_hi) {
  static BcEntry table[256];
  uint8_t kinds[256];
  memset(table, 0, sizeof(table));
  for (int i = 0; i < 256; ++i) {
    table[i].kind = bc_kind[i];
    table[i].rho = bc_rho[i];
    for (int a = 0; a < 3; ++a) table[i].u[a] = bc_u[i * 3 + a];
    kinds[i] = (uint8_t)bc_kind[i];
  }
  StepCall c;
  memset(&c, 0, sizeof(c));
  c.compute_dtype = compute_dtype;
  c.store_dtype = store_dtype;
  c.requested_v = cells_per_thread;
  c.f0 = f0;
  c.f1 = f1;
  c.bc = bc_mask;
  c.miss = missing_bits;
  c.table = table;
  c.table_rw = table;
  c.eq_omega_state = nullptr;
  c.kinds = kinds;
  c.nx = dims[0];
  c.ny = dims[1];
  c.nz = dims[2];
  c.x_begin = x_begin;
  c.x_count = x_count;
  c.omega = omega;
  c.ghost_lo = ghost_lo;
  c.ghost_hi = ghost_hi;
  c.out_lo = out_lo;
  c.out_hi = out_hi;
  static double eq_in[4][kMaxQ];  // as xlbn_stepper_create
  static uint8_t eq_ids[4];
  memset(eq_in, 0, sizeof(eq_in));
  memset(eq_ids, 0, sizeof(eq_ids));
  for (int id = 1, n = 0; id < 255 && n < 4; ++id) {
    if (table[id].kind != XLBN_BC_EQUILIBRIUM) continue;
    if (compute_dtype == XLBN_F32) equilibrium_on_host<MIRROR_LAT, float>(table[id].rho, table[id].u, eq_in[n]);
    else equilibrium_on_host<MIRROR_LAT, double>(table[id].rho, table[id].u, eq_in[n]);
    eq_ids[n++] = (uint8_t)id;
  }
  c.eq_in = &eq_in[0][0];
  c.eq_ids = eq_ids;
  for (int a = 0; a < 3; ++a) c.force[a] = force ? force[a] : 0.0;
  c.smagorinsky = smagorinsky;
  if (lattice == MIRROR_TAG && collision == (MIRROR_COLL)) return host_step_policy<MIRROR_LAT, (MIRROR_COLL)>(c);
  return fail(XLBN_E_ARG, "mirror: lattice %d / collision %d is not built", lattice, collision);
}

}  // extern "C"
