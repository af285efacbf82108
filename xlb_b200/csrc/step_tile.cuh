// The tile kernels: the fused step as a PERSISTENT, TMA-FED pipeline.  Same per-cell algebra, same results — bit for bit — as the
// direct-load kernels of step_kernel.cuh; what changes is how the bytes move.  Two kernels share the pipeline (producer warp, copy plan
// `tile_runs`, stage ring, mbarriers, direct stores):
//   step_tile_kernel<L, CELLS, CTAS>           FP32FP16 storage, BGK: two z-neighbours per thread as half2 words, packed fp32x2 arithmetic
//                                              (described first, below)
//   step_tile1_kernel<L, COLL, TC, TS, CTAS>   any storage type / collision, one cell per thread around the per-cell code of the direct
//                                              kernel; the default for D3Q19 BGK with fp32 storage — the headline path (second half of
//                                              this file)
//
// Why (written for the fp16 kernel, which came first).  With fp16 storage the step moves 77 B per cell and needs ~230 instructions per cell: the direct-load kernel keeps at most
// 32 warps/SM resident (64 registers, two cells per thread), every one of them parked on its own 29 loads (ncu: long_scoreboard 62 %,
// one eligible warp per cycle, 26 of 64 warp slots) — it stops at 0.71 (cavity) / 0.84 (periodic) of the HBM roofline, and 40 % of
// its integer work is 64-bit address arithmetic for 48 global accesses per thread.  Memory-level parallelism is tied to registers.
//
// Design (Blackwell: bulk asynchronous copies = TMA in 1-D form, mbarrier transaction counts, 227 KB of shared memory per SM).
//   * A TILE is 512 consecutive cells of one x-plane: R = 512 / nz whole rows (nz | 512).  For population l the tile's post-stream
//     values come from rows y - c_y of plane x - c_x: ONE contiguous run of 1 KB in global memory (two at the periodic y wrap), and
//     the c_z = +-1 shift is a one-element rotation inside each row — applied when the row is read back from shared memory.
//   * One PRODUCER warp per CTA: lane l issues population l's bulk copy (cp.async.bulk global -> shared, completion counted on the
//     stage's mbarrier), one more lane the tile's 512 bc ids.  A handful of address computations per TILE instead of per thread.
//     It runs ahead through a ring of input stages: the bytes in flight per SM are set by shared memory (2 CTAs x 3 stages x 19.5 KB
//     for D3Q19), not by registers or occupancy.
//   * 8 CONSUMER warps (256 threads x two z-neighbours): wait for the stage, read their half2 words with immediate-offset LDS (the
//     rotated rows take a second word and one PRMT), release the stage at once — a warp's arrival on the stage's "empty" mbarrier —
//     collide in packed fp32x2 and store the half2 results straight to global memory (coalesced 128 B per warp and population;
//     stores never stall a warp).  Consumer warps never meet at a CTA barrier: the first version staged the results in shared memory
//     for bulk stores and spent 28 % of its stall samples at that barrier (profiles/r2_ncu_tile_v1_*).
//   * Boundary cells.  FullwayBounceBack halves are a masked select of the opposite population's word, chosen per warp only when the
//     warp holds a boundary cell; EquilibriumBC halves are overwritten with the per-BC constant update (2-byte stores after the pair
//     store); any other kind, and cells with id 255 (never written), go through the scalar boundary routine of the direct kernel.
// Eligibility (xlbn_step falls back to the direct kernel otherwise): FP32FP16, BGK, no halo handle on the call (face planes of a slab
// store into peer memory: direct kernel), nz a divisor of 512 and a multiple of 8, ny a multiple of 512 / nz, 16-byte aligned arrays.
// Included by step_kernel.cuh (after the per-cell code it builds on, before the host-side launch code); not a stand-alone header.
#pragma once

namespace xlbn {

// CELLS = cells per tile (512: 8 consumer warps, two CTAs per SM; 1024: 16 consumer warps, one CTA per SM, 2 KB bulk copies where a
// population row is 1 KB — measured 0.93 against 0.87 of the roofline for the 512^3 cavity: half as many bulk copies per byte and one
// stage ring shared by twice as many warps).  The kernel takes it as a template argument; xlbn_step picks the largest that tiles the slab.
template <int CELLS>
struct TileDims {
  static constexpr int kCells = CELLS;
  static constexpr int kConsumers = CELLS / 2;       // consumer threads = half2 words per population row of a tile
  static constexpr int kThreads = kConsumers + 32;   // + the producer warp
  static constexpr int kRowBytes = CELLS * 2;
};
constexpr int kTileBarBytes = 128;
XLBN_DEV uint32_t tile_bits(__half2 h);
constexpr int kTileEqSlots = 4;  // EquilibriumBC ids whose constant update is kept in shared memory (more: those cells take the scalar routine)

// Per-CTA copy of the EquilibriumBC constants (BcEntry::eq_out, refreshed by bc_precompute_kernel when omega changes): slot_of[id] =
// slot of a boundary id or 0xff, word[slot][l] = the update of population l as a half2 word (e, e).
struct TileEqTable {
  uint32_t word[kTileEqSlots][kMaxQ];
  uint8_t slot_of[256];
};

template <class L>
XLBN_DEV void tile_eq_table_fill(const StepParams<__half>& p, TileEqTable& tab) {  // one thread
  int n = 0;
  for (int id = 0; id < 256; ++id) {
    tab.slot_of[id] = 0xff;
    if (id == 0 || id == 255 || p.kinds[id] != XLBN_BC_EQUILIBRIUM || n == kTileEqSlots) continue;
    tab.slot_of[id] = (uint8_t)n;
    for (int l = 0; l < L::Q; ++l) {
      const __half e = __float2half_rn(p.table[id].eq_out[l]);
      tab.word[n][l] = tile_bits(__halves2half2(e, e));
    }
    ++n;
  }
  for (; n < kTileEqSlots; ++n)
    for (int l = 0; l < L::Q; ++l) tab.word[n][l] = 0u;
}

// CTAS = resident CTAs per SM the kernel is compiled for (register budget 65536 / (288 * CTAS)); input stages fill what is left of
// the 227 KB of shared memory.
template <class L, int CELLS, int CTAS>
struct TileCfg {
  static constexpr int kInBytes = L::Q * TileDims<CELLS>::kRowBytes + CELLS;  // q population rows + the tile's bc ids
  static constexpr int kFixedBytes = kTileBarBytes + (int)((sizeof(TileEqTable) + 127) / 128 * 128);
  static constexpr int kMaxStages = (227 * 1024 / CTAS - 1024 - kFixedBytes) / kInBytes;
  static constexpr int kInStages = kMaxStages > 4 ? 4 : kMaxStages;
  static constexpr int kSmemBytes = kFixedBytes + kInStages * kInBytes;
  static_assert(L::Q + 1 <= 32, "one producer lane per population + one for the ids");
  static_assert(kInStages >= 2, "at least double buffering");
  static_assert(2 * kInStages * 8 <= 64 && kInStages * sizeof(unsigned) * 3 <= 64, "barriers and tile geometries share the first 128 bytes");
};

struct TileGeom {
  int x, y0;       // plane and first row of the tile
  unsigned cell0;  // element offset of its first cell inside a population
};

template <class TS>
XLBN_DEVFN inline TileGeom tile_geom(const StepParams<TS>& p, int tile, int rows, int tiles_per_plane) {
  TileGeom g;
  g.x = p.x_begin + tile / tiles_per_plane;
  g.y0 = (tile % tiles_per_plane) * rows;
  g.cell0 = (unsigned)g.x * (unsigned)p.plane + (unsigned)g.y0 * (unsigned)p.nz;
  return g;
}

// Where population l of a tile comes from: up to two contiguous runs (elements), in the order they are laid out in the stage row: run 0 =
// n0 elements from src0 to the start of the row, run 1 = n1 elements (0: none) from src1 right behind it.
// (cx, cy) = kernel-axis velocity components of population l: the producer lane looks them up once, before its tile loop.
// Scalars on purpose.  The first version filled src[2] / dst[2] / count[2] arrays by run-time index, and ptxas 12.9 merged dst[] and count[]
// in the scalar tile kernel's D3Q19 instantiation (wrong destination offsets on the wrapped rows; caught by the bit-identity tests).
template <class TS>
struct TileRuns {
  const TS *src0, *src1;
  unsigned n0, n1;
};

template <class L, class TS>
XLBN_DEVFN TileRuns<TS> tile_runs(const StepParams<TS>& p, int l, int cx, int cy, const TileGeom& g, int rows) {
  const int tab = (cx == 1 && g.x == 0) ? 1 : ((cx == -1 && g.x == p.nx - 1) ? 2 : 0);  // ghost plane / periodic wrap in x (fill_step_params)
  const TS* base = p.pull[tab][l] + (unsigned)g.x * (unsigned)p.plane;
  const unsigned nz = (unsigned)p.nz, all = (unsigned)rows * nz;
  const int ys = g.y0 - cy;  // first source row: pull from y - c_y (stream.py:66-78)
  TileRuns<TS> r;
  if (ys < 0) {  // row ny-1, then rows 0 .. rows-2
    r.src0 = base + (unsigned)(p.ny - 1) * nz;
    r.n0 = nz;
    r.src1 = base;
    r.n1 = all - nz;
  } else if (ys + rows > p.ny) {  // rows ys .. ny-1, then row 0
    r.src0 = base + (unsigned)ys * nz;
    r.n0 = all - nz;
    r.src1 = base;
    r.n1 = nz;
    if (r.n0 == 0) {  // a one-row tile: row 0 is all of it
      r.src0 = base;
      r.n0 = nz;
      r.n1 = 0;
    }
  } else {
    r.src0 = base + (unsigned)ys * nz;
    r.n0 = all;
    r.src1 = base;
    r.n1 = 0;
  }
  return r;
}

// two half2 words -> (hi half of a, lo half of b): the pair one element to the right of a / to the left of b
XLBN_DEV __half2 tile_funnel(uint32_t a, uint32_t b) {
  uint32_t r;
#if XLBN_ON_HOST
  r = (a >> 16) | (b << 16);
#else
  asm("prmt.b32 %0, %1, %2, 0x5432;" : "=r"(r) : "r"(a), "r"(b));
#endif
  union {
    uint32_t u;
    __half2 h;
  } c;
  c.u = r;
  return c.h;
}
XLBN_DEV uint32_t tile_bits(__half2 h) {
  union {
    __half2 h;
    uint32_t u;
  } c;
  c.h = h;
  return c.u;
}

// Thread t of a tile: its two cells' post-stream populations (as half2 words) and bc ids out of an input stage.
template <class L, int CELLS>
XLBN_DEV void tile_load(const StepParams<__half>& p, const uint32_t* in_words, const uint8_t* in_ids, const unsigned t, __half2 (&h)[L::Q], unsigned& ids) {
  constexpr int kTileConsumers = TileDims<CELLS>::kConsumers;
  const unsigned nz = (unsigned)p.nz, z0 = (2u * t) & (nz - 1u);  // nz is a power of two (a divisor of 512)
  const unsigned tp = (z0 == 0u) ? t + nz / 2u - 1u : t - 1u;          // word holding element z0 - 1 of the same row (periodic in z)
  const unsigned tn = (z0 + 2u == nz) ? t + 1u - nz / 2u : t + 1u;     // word holding element z0 + 2
  XLBN_FOR(L::Q, l)
    constexpr int cz = L::ck(2, l);
    const uint32_t w = in_words[l * kTileConsumers + t];
    if constexpr (cz == 0) {
      union {
        uint32_t u;
        __half2 h;
      } c;
      c.u = w;
      h[l] = c.h;
    } else if constexpr (cz == 1) {  // out[z] = in[z - 1]
      h[l] = tile_funnel(in_words[l * kTileConsumers + tp], w);
    } else {  // out[z] = in[z + 1]
      h[l] = tile_funnel(w, in_words[l * kTileConsumers + tn]);
    }
  XLBN_END
  ids = (unsigned)in_ids[2u * t] | ((unsigned)in_ids[2u * t + 1u] << 8);
}

// Collide the two cells and store; boundary cells as described at the top of the file.
template <class L, int CELLS>
XLBN_DEV void tile_compute(const StepParams<__half>& p, const TileEqTable& eq, const __half2 (&h)[L::Q], const unsigned ids, const unsigned t, const TileGeom& g) {
  using TS = __half;
  constexpr int Q = L::Q;
  const unsigned cell = g.cell0 + 2u * t;
  const float omega = (float)p.omega;
  const bool any_bc = ids != 0u;
  auto put = [&](auto l_, uint32_t word) {
    constexpr int l = decltype(l_)::value;
    union {
      uint32_t u;
      Pack<TS, 2> a;
    } c;
    c.u = word;
    gstore<TS, 2>(p.push[l] + cell, c.a);
  };
  if (!XLBN_ANY(0xffffffffu, any_bc)) {  // warp-uniform: no boundary cell in this warp
    h2_collide_each<L>(h, omega, [&](auto l_, f32x2 out) { put(l_, tile_bits(__float22half2_rn(out.v))); });
    return;
  }
  const int id0 = (int)(ids & 0xffu), id1 = (int)(ids >> 8);
  const int k0 = id0 ? (int)p.kinds[id0] : 0, k1 = id1 ? (int)p.kinds[id1] : 0;
  const bool eq0 = k0 == XLBN_BC_EQUILIBRIUM, eq1 = k1 == XLBN_BC_EQUILIBRIUM;
  const unsigned s0 = eq0 ? eq.slot_of[id0] : 0xffu, s1 = eq1 ? eq.slot_of[id1] : 0xffu;
  const bool complex_bc = (id0 == 255) | (id1 == 255) | (k0 != XLBN_BC_NONE && k0 != XLBN_BC_FULLWAY_BOUNCE_BACK && !eq0) |
                          (k1 != XLBN_BC_NONE && k1 != XLBN_BC_FULLWAY_BOUNCE_BACK && !eq1) | (eq0 && s0 == 0xffu) | (eq1 && s1 == 0xffu) |
                          (eq0 && eq1 && s0 != s1);
  // FullwayBounceBack halves: out[l] = f_post_stream[opp l], a bit copy (bc_fullway_bounce_back.py:60-72).
  // EquilibriumBC halves: bc_equilibrium.py:76-86 followed by the ordinary collision = a per-BC constant, kept in shared memory.
  const uint32_t m_fw = (k0 == XLBN_BC_FULLWAY_BOUNCE_BACK ? 0x0000ffffu : 0u) | (k1 == XLBN_BC_FULLWAY_BOUNCE_BACK ? 0xffff0000u : 0u);
  const uint32_t m_eq = complex_bc ? 0u : ((eq0 ? 0x0000ffffu : 0u) | (eq1 ? 0xffff0000u : 0u));
  if (!XLBN_ANY(0xffffffffu, m_eq != 0u)) {  // warp-uniform: walls only
    h2_collide_each<L>(h, omega, [&](auto l_, f32x2 out) {
      constexpr int l = decltype(l_)::value;
      const uint32_t o = tile_bits(__float22half2_rn(out.v));
      if (!complex_bc) put(l_, (o & ~m_fw) | (tile_bits(h[L::opp(l)]) & m_fw));
    });
  } else {
    const uint32_t* ew = eq.word[eq0 ? s0 : (eq1 ? s1 : 0u)];  // lanes without an EquilibriumBC cell read slot 0 and mask it away
    if (complex_bc) ew = eq.word[0];
    h2_collide_each<L>(h, omega, [&](auto l_, f32x2 out) {
      constexpr int l = decltype(l_)::value;
      const uint32_t o = tile_bits(__float22half2_rn(out.v));
      if (!complex_bc) put(l_, (o & ~(m_fw | m_eq)) | (tile_bits(h[L::opp(l)]) & m_fw) | (ew[l] & m_eq));
    });
  }
  if (!complex_bc) return;
  // every other kind, and cells with id 255 (not written: nse_stepper.py:356-358): the scalar boundary routine of the direct kernel
  const unsigned nz = (unsigned)p.nz, j = 2u * t;
  const int y = g.y0 + (int)(j / nz), z0 = (int)(j & (nz - 1u));
  float fs[2][Q];
  XLBN_FOR(Q, l)
    const float2 f = __half22float2(h[l]);
    fs[0][l] = f.x;
    fs[1][l] = f.y;
  XLBN_END
  Pack<uint8_t, 2> pk;
  pk.v[0] = (uint8_t)id0;
  pk.v[1] = (uint8_t)id1;
  bc_compute<L, XLBN_BGK, float, TS, 2>(p, pk, g.x, y, z0, fs);
  if ((id0 == 255) | (id1 == 255)) store_cells<L, float, TS, 2, 0, true>(p, cell, pk, fs);
  else store_cells<L, float, TS, 2, 0, false>(p, cell, pk, fs);
}

template <class L, int CELLS>
bool tile_eligible(const StepParams<__half>& p, const void* f0, const void* f1, const void* ghost_lo, const void* ghost_hi, bool has_peers) {
  constexpr int kTileCells = CELLS;
  auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  if (has_peers || p.nz < 8 || p.nz > kTileCells || (kTileCells % p.nz) != 0 || (p.nz % 8) != 0) return false;
  const int rows = kTileCells / p.nz;
  if (p.ny % rows != 0) return false;
  if (((long long)p.plane * 2) % 16 != 0) return false;  // ghost planes and population strides keep the 16-byte alignment
  return aligned(f0) && aligned(f1) && aligned(ghost_lo) && aligned(ghost_hi) && aligned(p.bc);
}

// ---- the scalar tile kernel's per-thread code (the kernel itself is further down; tests/host_math/mirror_step.cu runs these on the host) ----
constexpr int kT1Cells = 512;

// One cell out of a stage: population l of cell t is element t - c_z (inside its row) of stage row l.
template <class L, class TC, class TS>
XLBN_DEV void tile1_load(const TS* in, const unsigned char* id_row, unsigned t, unsigned tm, unsigned tp, TC (&f)[1][L::Q], Pack<uint8_t, 1>& ids) {
  XLBN_FOR(L::Q, l)
    constexpr int cz = L::ck(2, l);
    f[0][l] = Cvt<TC, TS>::up(in[l * kT1Cells + (cz == 1 ? tm : (cz == -1 ? tp : t))]);
  XLBN_END
  ids.v[0] = id_row[t];
}

// ... and from there on the per-cell code of the direct kernel (step_body): input-side EquilibriumBC, collision, boundary routine, store.
template <class L, int COLL, class TC, class TS>
XLBN_DEV void tile1_finish(const StepParams<TS>& p, const TileGeom& g, unsigned t, unsigned z, TC omega, TC (&f)[1][L::Q], const Pack<uint8_t, 1>& ids) {
  const int id = ids.v[0];
  if (id == 255) return;  // nse_stepper.py:356-358
  const unsigned cell = g.cell0 + t;
  bool tail = id != 0;
  if (tail) {  // EquilibriumBC at the input side (StepParams::eq_in)
    int slot = -1;
#pragma unroll
    for (int i = 0; i < kEqSlots; ++i)
      if (p.eq_ids[i] == id) slot = i;
    if (slot >= 0) {
      XLBN_FOR(L::Q, l) f[0][l] = (TC)p.eq_in[slot][l]; XLBN_END
      tail = false;
    }
  }
  if (!tail) {  // the straight-line path ends here, so that its register allocation is independent of the boundary code (as in step_body)
    collide_in_step<L, COLL, TC, TS>(p, f[0], omega);
    store_cells<L, TC, TS, 1, 0, false>(p, cell, ids, f);
    return;
  }
  const int y = g.y0 + (int)(t / (unsigned)p.nz);
  bc_compute<L, COLL, TC, TS, 1>(p, ids, g.x, y, (int)z, f);
  store_cells<L, TC, TS, 1, 0, false>(p, cell, ids, f);
}

template <class L, class TS>
bool tile1_eligible(const StepParams<TS>& p, const void* f0, const void* f1, const void* ghost_lo, const void* ghost_hi, bool has_peers) {
  auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % 16) == 0; };
  if (has_peers || p.nz < 8 || p.nz > kT1Cells || (kT1Cells % p.nz) != 0 || ((p.nz * (int)sizeof(TS)) % 16) != 0 || (p.nz % 16) != 0) return false;
  if (p.ny % (kT1Cells / p.nz) != 0) return false;
  return aligned(f0) && aligned(f1) && aligned(ghost_lo) && aligned(ghost_hi) && aligned(p.bc);
}

#if !XLBN_ON_HOST
// ---- device-only plumbing: mbarrier, bulk copies, named barrier -------------------------------------------------------------------
namespace tile_ptx {
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(saddr(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst_smem)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}
}  // namespace tile_ptx

template <class L, int CELLS, int CTAS>
__global__ void __launch_bounds__(TileDims<CELLS>::kThreads, CTAS) step_tile_kernel(const __grid_constant__ StepParams<__half> p, const int n_tiles, const int rows, const int tiles_per_plane) {
  using namespace tile_ptx;
  using C = TileCfg<L, CELLS, CTAS>;
  constexpr int kTileCells = CELLS, kTileConsumers = TileDims<CELLS>::kConsumers, kTileRowBytes = TileDims<CELLS>::kRowBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // one per input stage: the bulk copies of the stage have landed
  uint64_t* empty = full + C::kInStages;                // one per input stage: all 8 consumer warps have taken their words
  TileGeom* geoms = reinterpret_cast<TileGeom*>(smem + 64);  // per input stage: where the tile in it lives (written by the producer)
  TileEqTable& eq = *reinterpret_cast<TileEqTable*>(smem + kTileBarBytes);
  unsigned char* in0 = smem + C::kFixedBytes;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::kInStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kTileConsumers / 32);
    }
    mbar_init_fence();
  }
  if (tid == 32) tile_eq_table_fill<L>(p, eq);
  __syncthreads();

  if (warp == kTileConsumers / 32) {
    // ---- producer: runs ahead of the consumers by up to kInStages tiles ----
    const int pl = lane < L::Q ? lane : 0, pcx = L::ck(0, pl), pcy = L::ck(1, pl);
    int k = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++k) {
      const int s = k % C::kInStages;
      const uint32_t phase = (uint32_t)(k / C::kInStages) & 1u;
      mbar_wait(empty + s, phase ^ 1u);  // first round: passes at once
      const TileGeom g = tile_geom(p, tile, rows, tiles_per_plane);
      unsigned char* stage = in0 + s * C::kInBytes;
      if (lane == 0) {
        geoms[s] = g;  // the consumers read it after the stage's barrier: the 2 integer divisions of tile_geom are done once per tile
        mbar_expect_tx(full + s, (uint32_t)C::kInBytes);
      }
      __syncwarp();
      if (lane < L::Q) {
        const TileRuns<__half> r = tile_runs<L, __half>(p, lane, pcx, pcy, g, rows);
        unsigned char* row = stage + lane * kTileRowBytes;
        bulk_load(row, r.src0, r.n0 * 2u, full + s);
        if (r.n1) bulk_load(row + r.n0 * 2u, r.src1, r.n1 * 2u, full + s);
      } else if (lane == L::Q) {
        bulk_load(stage + L::Q * kTileRowBytes, p.bc + g.cell0, kTileCells, full + s);
      }
    }
    return;
  }

  // ---- consumers: 8 independent warps, no CTA-wide synchronisation ----
  int k = 0;
  for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++k) {
    const int s = k % C::kInStages;
    const uint32_t phase = (uint32_t)(k / C::kInStages) & 1u;
    const unsigned char* stage = in0 + s * C::kInBytes;
    // which 32 words a warp takes rotates from tile to tile: in a closed box the floor / lid cells sit at the two ends of EVERY row, and
    // the stage ring advances at the pace of the slowest warp — so every warp takes its turn with them
    const unsigned t = ((unsigned)tid + 32u * (unsigned)(k & (kTileConsumers / 32 - 1))) & (unsigned)(kTileConsumers - 1);
    __half2 h[L::Q];
    unsigned ids;
    mbar_wait(full + s, phase);
    const TileGeom g = geoms[s];
    tile_load<L, CELLS>(p, reinterpret_cast<const uint32_t*>(stage), stage + L::Q * kTileRowBytes, t, h, ids);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);  // the stage can be refilled while this tile is computed
    tile_compute<L, CELLS>(p, eq, h, ids, t, g);
  }
}

template <class L, int CELLS, int CTAS>
int launch_step_tile(const StepParams<__half>& p, int x_count, cudaStream_t stream) {
  using C = TileCfg<L, CELLS, CTAS>;
  constexpr int kTileCells = CELLS, kTileThreads = TileDims<CELLS>::kThreads;
  static int sm_counts[64] = {0};  // per device: the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  int dev = 0;
  XLBN_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(XLBN_E_STATE, "tile kernel: device ordinal %d", dev);
  if (sm_counts[dev] == 0) {
    int n = 0;
    XLBN_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    XLBN_CUDA_OK(cudaFuncSetAttribute(step_tile_kernel<L, CELLS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    sm_counts[dev] = n;
  }
  const int sm_count = sm_counts[dev];
  const int rows = kTileCells / p.nz, tiles_per_plane = p.ny / rows;
  const long long n_tiles = (long long)tiles_per_plane * x_count;
  if (n_tiles > 0x7fffffffLL) return fail(XLBN_E_SHAPE, "tile kernel: %lld tiles", n_tiles);
  const long long resident = (long long)CTAS * sm_count;
  const int grid = (int)(n_tiles < resident ? n_tiles : resident);  // persistent: CTAS CTAs per SM walk the tiles round-robin
  step_tile_kernel<L, CELLS, CTAS><<<grid, kTileThreads, C::kSmemBytes, stream>>>(p, (int)n_tiles, rows, tiles_per_plane);
  XLBN_LAUNCH_OK("step_tile_kernel launch");
  return 0;
}

// ---- the scalar tile kernel: ONE cell per consumer thread, any collision, any precision policy -------------------------------------
// The same pipeline (producer warp, bulk copies into a ring of stages, 16 consumer warps that never meet at a CTA barrier, direct
// stores) around the per-cell code of the direct kernel: the pulled populations come out of shared memory instead of 19 / 27 global loads
// with 64-bit address arithmetic, everything after that is step_body's — input-side EquilibriumBC, collide_in_step, the scalar boundary
// routine.  A tile is 512 cells; z rotation = one shifted LDS per population.  Persistent, so a small grid has no tail of partial waves.
constexpr int kT1Threads = kT1Cells + 32;

template <class L, class TS, int CTAS>
struct Tile1Cfg {
  static constexpr int kInBytes = L::Q * kT1Cells * (int)sizeof(TS) + kT1Cells;
  static constexpr int kMaxStages = (227 * 1024 / CTAS - 1024 - kTileBarBytes) / kInBytes;
  static constexpr int kInStages = kMaxStages > 4 ? 4 : kMaxStages;
  static constexpr int kSmemBytes = kTileBarBytes + kInStages * kInBytes;
  static_assert(kInStages >= 2, "at least double buffering");
};

template <class L, int COLL, class TC, class TS, int CTAS>
__global__ void __launch_bounds__(kT1Threads, CTAS) step_tile1_kernel(const __grid_constant__ StepParams<TS> p, const int n_tiles, const int rows, const int tiles_per_plane) {
  using namespace tile_ptx;
  using C = Tile1Cfg<L, TS, CTAS>;
  constexpr int Q = L::Q;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + C::kInStages;
  TileGeom* geoms = reinterpret_cast<TileGeom*>(smem + 64);
  unsigned char* in0 = smem + kTileBarBytes;
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < C::kInStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kT1Cells / 32);
    }
    mbar_init_fence();
  }
  __syncthreads();

  if (warp == kT1Cells / 32) {  // producer
    const int pl = lane < Q ? lane : 0, pcx = L::ck(0, pl), pcy = L::ck(1, pl);
    int k = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++k) {
      const int s = k % C::kInStages;
      const uint32_t phase = (uint32_t)(k / C::kInStages) & 1u;
      mbar_wait(empty + s, phase ^ 1u);
      const TileGeom g = tile_geom(p, tile, rows, tiles_per_plane);
      unsigned char* stage = in0 + s * C::kInBytes;
      if (lane == 0) {
        geoms[s] = g;
        mbar_expect_tx(full + s, (uint32_t)C::kInBytes);
      }
      __syncwarp();
      if (lane < Q) {
        const TileRuns<TS> r = tile_runs<L, TS>(p, lane, pcx, pcy, g, rows);
        unsigned char* row = stage + (unsigned)lane * (unsigned)(kT1Cells * sizeof(TS));
        bulk_load(row, r.src0, r.n0 * (unsigned)sizeof(TS), full + s);
        if (r.n1) bulk_load(row + r.n0 * (unsigned)sizeof(TS), r.src1, r.n1 * (unsigned)sizeof(TS), full + s);
      } else if (lane == Q) {
        bulk_load(stage + Q * kT1Cells * sizeof(TS), p.bc + g.cell0, kT1Cells, full + s);
      }
    }
    return;
  }

  // consumers
  const unsigned nz = (unsigned)p.nz;
  const TC omega = (TC)p.omega;
  int k = 0;
  for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++k) {
    const int s = k % C::kInStages;
    const uint32_t phase = (uint32_t)(k / C::kInStages) & 1u;
    const unsigned char* stage = in0 + s * C::kInBytes;
    const TS* in = reinterpret_cast<const TS*>(stage);
    const unsigned t = ((unsigned)tid + 32u * (unsigned)(k & (kT1Cells / 32 - 1))) & (unsigned)(kT1Cells - 1);  // warps take turns with the row ends
    const unsigned z = t & (nz - 1u);
    const unsigned tm = (z == 0u) ? t + nz - 1u : t - 1u;       // source of c_z = +1 populations: element z - 1 of the same row
    const unsigned tp = (z + 1u == nz) ? t + 1u - nz : t + 1u;  // source of c_z = -1 populations
    mbar_wait(full + s, phase);
    const TileGeom g = geoms[s];
    TC f[1][Q];
    Pack<uint8_t, 1> ids;
    tile1_load<L, TC, TS>(in, stage + Q * kT1Cells * sizeof(TS), t, tm, tp, f, ids);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    tile1_finish<L, COLL, TC, TS>(p, g, t, z, omega, f, ids);
  }
}

template <class L, int COLL, class TC, class TS, int CTAS>
int launch_step_tile1(const StepParams<TS>& p, int x_count, cudaStream_t stream) {
  using C = Tile1Cfg<L, TS, CTAS>;
  static int sm_counts[64] = {0};
  int dev = 0;
  XLBN_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(XLBN_E_STATE, "tile kernel: device ordinal %d", dev);
  if (sm_counts[dev] == 0) {
    int n = 0;
    XLBN_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    XLBN_CUDA_OK(cudaFuncSetAttribute(step_tile1_kernel<L, COLL, TC, TS, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    sm_counts[dev] = n;
  }
  const int rows = kT1Cells / p.nz, tiles_per_plane = p.ny / rows;
  const long long n_tiles = (long long)tiles_per_plane * x_count;
  if (n_tiles > 0x7fffffffLL) return fail(XLBN_E_SHAPE, "tile kernel: %lld tiles", n_tiles);
  const long long resident = (long long)CTAS * sm_counts[dev];
  const int grid = (int)(n_tiles < resident ? n_tiles : resident);
  step_tile1_kernel<L, COLL, TC, TS, CTAS><<<grid, kT1Threads, C::kSmemBytes, stream>>>(p, (int)n_tiles, rows, tiles_per_plane);
  XLBN_LAUNCH_OK("step_tile1_kernel launch");
  return 0;
}
#endif  // !XLBN_ON_HOST

}  // namespace xlbn
