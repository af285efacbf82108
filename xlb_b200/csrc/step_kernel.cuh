// The hot path: ONE fused kernel per time step — pull-stream + streaming BCs + macroscopic + equilibrium + collision +
// collision BCs + aux recovery + store.  Replaces the reference's Warp kernel xlb/operator/stepper/nse_stepper.py:344-381
// (and its JAX twin 147-192).
//
// B200 design:
//  * layout [q][nx][ny][nz], z unit-stride.  threadIdx.x runs along z; every thread owns V CONSECUTIVE z-cells so that
//    each population is moved with one 8/16-byte load and one 8/16-byte store per thread (fully coalesced, sector
//    aligned).  Populations with c_z = ±1 need the row shifted by one element: they take the same aligned vector load
//    plus ONE scalar load of the element just outside the vector (same 128-B lines -> L1 hits; DRAM traffic stays
//    2*q*sizeof(store)+1 B/cell).  x/y neighbours are whole-row offsets and stay aligned.
//  * every population element is read by exactly one thread (the pull is a bijection), so there is no reuse to stage in
//    shared memory; occupancy and bytes in flight (q*V*sizeof(store) per thread) are what feed HBM.
//  * store<->compute conversion is fused into the load/store (PrecisionPolicy), all algebra is in registers.
//  * bc_mask is 1 B/cell; id 0 takes the straight-line path, 255 is skipped, anything else calls a non-inlined
//    boundary-cell routine that alone touches the missing-direction bitmask, the cell's own pre-stream populations and
//    the aux value (O(N^2) cells).  The reference re-reads f_1 and a q-byte missing mask for EVERY cell
//    (nse_stepper.py:296-316).
//  * x-slab multi-GPU: pulls across the slab faces read compact ghost planes; the kernel that updates planes 0 / nx-1
//    also stores the outgoing populations straight into the neighbour GPUs' ghost planes through peer-mapped pointers
//    (compute and NVLink transfer in the same kernel; no pack / exchange pass).
#pragma once

#include "lbm_math.cuh"

namespace xlbn {

struct BcEntry {
  int kind;
  int pad;
  double rho;
  double u[3];
};

template <class TS>
struct StepParams {
  const TS* f0;
  TS* f1;
  TS* f0w;  // writable alias of f0: aux recovery only (nse_stepper.py:338)
  const uint8_t* bc;
  const uint32_t* miss;
  const BcEntry* table;  // 256 entries, indexed by bc id
  int nx, ny, nz, x_begin;
  long long plane;  // ny * nz
  long long n;      // nx * ny * nz  (population stride)
  double omega;
  const TS* ghost_lo;  // plane "x = -1":  populations with ck(0) = +1, [n_xdir][ny][nz]; NULL -> periodic wrap
  const TS* ghost_hi;  // plane "x = nx":  populations with ck(0) = -1
  TS* out_lo;          // lo neighbour's ghost_hi for the NEXT step (peer memory) or NULL
  TS* out_hi;          // hi neighbour's ghost_lo for the NEXT step (peer memory) or NULL
};

// Start of x-plane (x - cx) of population l as seen by a pull: inside the array, in a ghost plane, or wrapped.
template <class L, class TS, int l>
XLBN_DEV const TS* pull_plane(const StepParams<TS>& p, int x) {
  constexpr int cx = L::ck(0, l);
  const TS* base = p.f0 + (long long)l * p.n;
  if constexpr (cx == 0) {
    return base + (long long)x * p.plane;
  } else if constexpr (cx == 1) {
    if (x > 0) return base + (long long)(x - 1) * p.plane;
    return p.ghost_lo ? p.ghost_lo + (long long)L::xdir_slot(l) * p.plane : base + (long long)(p.nx - 1) * p.plane;
  } else {
    if (x < p.nx - 1) return base + (long long)(x + 1) * p.plane;
    return p.ghost_hi ? p.ghost_hi + (long long)L::xdir_slot(l) * p.plane : base;
  }
}

// f0[l] at an arbitrary (possibly out-of-range) kernel-coordinate cell: periodic in y/z; x through ghost or wrap.
template <class L, class TC, class TS>
__device__ TC load_f0_any(const StepParams<TS>& p, int l, int ck0, int slot, int x, int y, int z) {
  y = (y % p.ny + p.ny) % p.ny;
  z = (z % p.nz + p.nz) % p.nz;
  const long long yz = (long long)y * p.nz + z;
  if (x == -1 && p.ghost_lo && ck0 == 1) return Cvt<TC, TS>::up(p.ghost_lo[(long long)slot * p.plane + yz]);
  if (x == p.nx && p.ghost_hi && ck0 == -1) return Cvt<TC, TS>::up(p.ghost_hi[(long long)slot * p.plane + yz]);
  x = (x % p.nx + p.nx) % p.nx;
  return Cvt<TC, TS>::up(p.f0[(long long)l * p.n + (long long)x * p.plane + yz]);
}

// Full update of ONE boundary cell: fio holds the pulled (post-stream) populations on entry and the values to store
// on exit.  Order of operations = reference kernel: streaming BC -> collide -> collision BC / outflow aux -> aux recovery
// (nse_stepper.py:361-381).
template <class L, int COLL, class TC, class TS>
__device__ __noinline__ void bc_cell(const StepParams<TS>& p, int id, int x, int y, int z, TC* fio) {
  constexpr int Q = L::Q;
  TC f[Q];
  XLBN_FOR(Q, l) f[l] = fio[l]; XLBN_END
  const BcEntry e = p.table[id];
  const int kind = e.kind;
  const long long cell = (long long)x * p.plane + (long long)y * p.nz + z;
  const uint32_t miss = (bc_kind_needs_missing(kind) && p.miss) ? p.miss[cell] : 0u;
  const TC omega = (TC)p.omega;

  if (kind == XLBN_BC_FULLWAY_BOUNCE_BACK) {
    // collision-step BC: the collided value is discarded, out[l] = f_post_stream[opp[l]] (bc_fullway_bounce_back.py:60-72)
    XLBN_FOR(Q, l) fio[l] = f[L::opp(l)]; XLBN_END
    return;
  }

  if (kind == XLBN_BC_EQUILIBRIUM) {
    TC u[L::D];
    XLBN_FOR(L::D, d) u[d] = (TC)e.u[d]; XLBN_END
    equilibrium<L, TC>((TC)e.rho, u, f);  // bc_equilibrium.py:76-86
  } else if (bc_kind_needs_fpre(kind)) {
    TC fpre[Q];  // the cell's own PRE-stream populations (nse_stepper.py:363-367)
    XLBN_FOR(Q, l) fpre[l] = Cvt<TC, TS>::up(p.f0[(long long)l * p.n + cell]); XLBN_END
    if (kind == XLBN_BC_DO_NOTHING) {
      XLBN_FOR(Q, l) f[l] = fpre[l]; XLBN_END  // bc_do_nothing.py:52-63
    } else {
      bc_take_opposite_of_pre<L, TC>(fpre, miss, f);
    }
  } else if (bc_kind_needs_aux(kind)) {
    // prescribed value lives in f1[0, cell] (boundary_condition.py:151) and is handed back through f0[0, cell] so that
    // it survives the caller's buffer swap (nse_stepper.py:318-342)
    const TS raw = p.f1[cell];
    bc_zouhe<L, TC>(kind, Cvt<TC, TS>::up(raw), miss, f);
    p.f0w[cell] = raw;
  }

  if (kind == XLBN_BC_EXTRAPOLATION_OUTFLOW) {
    // post-collision aux update (bc_extrapolation_outflow.py:172-195): for missing l
    //   out[opp[l]] = (1 - cs) f_post_stream[l] + cs f0[l, cell - (c_l + n)]
    int ni[L::D];
    bc_normal<L>(miss, ni);
    int nk[3] = {0, 0, 0};
    XLBN_FOR(L::D, d) nk[d + 3 - L::D] = ni[d]; XLBN_END
    const TC cs = TC(0.57735026918962576451);
    TC aux[Q];
    XLBN_FOR(Q, l)
      aux[l] = TC(0);
      if ((miss >> l) & 1u) {
        const TC fn = load_f0_any<L, TC, TS>(p, l, L::ck(0, l), L::ck(0, l) != 0 ? L::xdir_slot(l) : 0, x - (L::ck(0, l) + nk[0]),
                                             y - (L::ck(1, l) + nk[1]), z - (L::ck(2, l) + nk[2]));
        aux[l] = (TC(1.0) - cs) * f[l] + cs * fn;
      }
    XLBN_END
    collide_cell<L, COLL, TC>(f, omega);
    XLBN_FOR(Q, l)
      if ((miss >> l) & 1u) f[L::opp(l)] = aux[l];
    XLBN_END
  } else {
    collide_cell<L, COLL, TC>(f, omega);
  }
  XLBN_FOR(Q, l) fio[l] = f[l]; XLBN_END
}

template <class L, int COLL, class TC, class TS, int V>
struct StepTraits {
  static constexpr int kThreads = 128;
  // register estimate: V*Q population registers (x2 for fp64) + algebra temporaries
  static constexpr int kRegs = V * L::Q * (int)(sizeof(TC) / 4) + (COLL == XLBN_KBC ? 3 * L::Q * (int)(sizeof(TC) / 4) : 40) + 24;
  static constexpr int kMinBlocksRaw = 65536 / (kThreads * (kRegs > 255 ? 255 : kRegs));
  static constexpr int kMinBlocks = kMinBlocksRaw < 1 ? 1 : (kMinBlocksRaw > 12 ? 12 : kMinBlocksRaw);
};

template <class L, int COLL, class TC, class TS, int V>
__global__ void __launch_bounds__(StepTraits<L, COLL, TC, TS, V>::kThreads, StepTraits<L, COLL, TC, TS, V>::kMinBlocks)
    step_kernel(const __grid_constant__ StepParams<TS> p) {
  constexpr int Q = L::Q;
  const int zv = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = p.x_begin + blockIdx.z;
  const int z0 = zv * V;
  if (z0 >= p.nz || y >= p.ny) return;

  const int nz = p.nz;
  const long long cell = (long long)x * p.plane + (long long)y * nz + z0;

  // boundary ids of the V cells (one V-byte load)
  const Pack<uint8_t, V> ids = load_pack<uint8_t, V>(p.bc + cell);
  bool any_solid = false, all_solid = true;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    any_solid |= (ids.v[v] == 255);
    all_solid &= (ids.v[v] == 255);
  }
  if (all_solid) return;  // nse_stepper.py:356-358

  // rows the pull reads from: y - cy with periodic wrap (stream.py:66-78)
  const long long row_c = (long long)y * nz;
  const long long row_m = (long long)(y == 0 ? p.ny - 1 : y - 1) * nz;   // source row for cy = +1
  const long long row_p = (long long)(y == p.ny - 1 ? 0 : y + 1) * nz;   // source row for cy = -1
  const int z_lo = (z0 == 0) ? nz - 1 : z0 - 1;                           // element left of the vector  (cz = +1)
  const int z_hi = (z0 + V >= nz) ? 0 : z0 + V;                           // element right of the vector (cz = -1)

  TC f[V][Q];
  XLBN_FOR(Q, l)
    constexpr int cy = L::ck(1, l), cz = L::ck(2, l);
    const TS* row = pull_plane<L, TS, l>(p, x) + (cy == 1 ? row_m : (cy == -1 ? row_p : row_c));
    if constexpr (cz == 0) {
      const Pack<TS, V> a = load_pack<TS, V>(row + z0);
#pragma unroll
      for (int v = 0; v < V; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v]);
    } else if constexpr (V == 1) {
      f[0][l] = Cvt<TC, TS>::up(row[cz == 1 ? z_lo : z_hi]);
    } else if constexpr (cz == 1) {  // out[z] = in[z - 1]
      const Pack<TS, V> a = load_pack<TS, V>(row + z0);
      const TS e = row[z_lo];
      f[0][l] = Cvt<TC, TS>::up(e);
#pragma unroll
      for (int v = 1; v < V; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v - 1]);
    } else {  // out[z] = in[z + 1]
      const Pack<TS, V> a = load_pack<TS, V>(row + z0);
      const TS e = row[z_hi];
#pragma unroll
      for (int v = 0; v < V - 1; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v + 1]);
      f[V - 1][l] = Cvt<TC, TS>::up(e);
    }
  XLBN_END

  const TC omega = (TC)p.omega;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int id = ids.v[v];
    if (id == 0) {
      collide_cell<L, COLL, TC>(f[v], omega);
    } else if (id != 255) {
      TC tmp[Q];
      XLBN_FOR(Q, l) tmp[l] = f[v][l]; XLBN_END
      bc_cell<L, COLL, TC, TS>(p, id, x, y, z0 + v, tmp);
      XLBN_FOR(Q, l) f[v][l] = tmp[l]; XLBN_END
    }
  }

  // store (fused compute -> store conversion); outgoing face populations also go to the neighbour GPUs' ghost planes
  const bool to_hi = (p.out_hi != nullptr) && (x == p.nx - 1);
  const bool to_lo = (p.out_lo != nullptr) && (x == 0);
  const long long yz = row_c + z0;
  if (!any_solid) {
    XLBN_FOR(Q, l)
      Pack<TS, V> a;
#pragma unroll
      for (int v = 0; v < V; ++v) a.v[v] = Cvt<TC, TS>::down(f[v][l]);
      store_pack<TS, V>(p.f1 + (long long)l * p.n + cell, a);
      if constexpr (L::ck(0, l) == 1) {
        if (to_hi) store_pack<TS, V>(p.out_hi + (long long)L::xdir_slot(l) * p.plane + yz, a);
      } else if constexpr (L::ck(0, l) == -1) {
        if (to_lo) store_pack<TS, V>(p.out_lo + (long long)L::xdir_slot(l) * p.plane + yz, a);
      }
    XLBN_END
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (ids.v[v] == 255) continue;
      XLBN_FOR(Q, l)
        const TS s = Cvt<TC, TS>::down(f[v][l]);
        p.f1[(long long)l * p.n + cell + v] = s;
        if constexpr (L::ck(0, l) == 1) {
          if (to_hi) p.out_hi[(long long)L::xdir_slot(l) * p.plane + yz + v] = s;
        } else if constexpr (L::ck(0, l) == -1) {
          if (to_lo) p.out_lo[(long long)L::xdir_slot(l) * p.plane + yz + v] = s;
        }
      XLBN_END
    }
  }
}

// ---- host-side launch ------------------------------------------------------------------------------------------------
template <class L, int COLL, class TC, class TS, int V>
int launch_step_v(const StepParams<TS>& p, int x_count, cudaStream_t stream) {
  constexpr int T = StepTraits<L, COLL, TC, TS, V>::kThreads;
  const int nzv = (p.nz + V - 1) / V;
  int bx = 32;
  while (bx < nzv && bx < T) bx *= 2;
  const int by = T / bx;
  dim3 block(bx, by, 1);
  dim3 grid((nzv + bx - 1) / bx, (p.ny + by - 1) / by, x_count);
  if (grid.y > 65535u || grid.z > 65535u) return fail(XLBN_E_SHAPE, "grid too large for launch: ny=%d x_count=%d", p.ny, x_count);
  step_kernel<L, COLL, TC, TS, V><<<grid, block, 0, stream>>>(p);
  XLBN_LAUNCH_OK("step_kernel launch");
  return 0;
}

// V must divide nz and keep every vector access aligned; otherwise fall back to the next smaller V.
template <class TS>
int pick_cells_per_thread(int requested, int dflt, const StepParams<TS>& p) {
  int v = requested > 0 ? requested : dflt;
  const int vmax = 16 / (int)sizeof(TS);
  if (v > vmax) v = vmax;
  while (v & (v - 1)) --v;  // power of two
  auto aligned = [&](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  while (v > 1) {
    const size_t a = sizeof(TS) * v;
    if (p.nz % v == 0 && aligned(p.f0, a) && aligned(p.f1, a) && aligned(p.ghost_lo, a) && aligned(p.ghost_hi, a) &&
        aligned(p.out_lo, a) && aligned(p.out_hi, a) && aligned(p.bc, v))
      break;
    v /= 2;
  }
  return v;
}

template <class L, int COLL, class TC, class TS>
int launch_step(const StepParams<TS>& p, int x_count, int requested_v, cudaStream_t stream) {
  // defaults chosen on B200 (profiles/): 16-byte accesses for D3Q19, 8-byte for D3Q27 (register budget)
  constexpr int dflt = (sizeof(TS) == 8) ? 2 : ((L::Q > 19) ? 2 : 4);
  const int v = pick_cells_per_thread<TS>(requested_v, dflt, p);
  switch (v) {
    case 1: return launch_step_v<L, COLL, TC, TS, 1>(p, x_count, stream);
    case 2: return launch_step_v<L, COLL, TC, TS, 2>(p, x_count, stream);
    case 4:
      if constexpr (sizeof(TS) <= 4) return launch_step_v<L, COLL, TC, TS, 4>(p, x_count, stream);
    case 8:
      if constexpr (sizeof(TS) <= 2) return launch_step_v<L, COLL, TC, TS, 8>(p, x_count, stream);
  }
  return fail(XLBN_E_ARG, "cells_per_thread = %d not available for this store dtype", v);
}

// One entry per (lattice, collision); dispatches on (compute, store) dtype.  Defined in step_inst_*.cu.
struct StepCall {
  int compute_dtype, store_dtype, requested_v;
  const void* f0;
  void* f1;
  const uint8_t* bc;
  const uint32_t* miss;
  const BcEntry* table;
  int nx, ny, nz, x_begin, x_count;
  double omega;
  const void* ghost_lo;
  const void* ghost_hi;
  void* out_lo;
  void* out_hi;
  cudaStream_t stream;
};

template <class L, int COLL>
int dispatch_step(const StepCall& c);

template <class L, int COLL, class TC, class TS>
int run_step_typed(const StepCall& c) {
  StepParams<TS> p;
  p.f0 = static_cast<const TS*>(c.f0);
  p.f1 = static_cast<TS*>(c.f1);
  p.f0w = const_cast<TS*>(static_cast<const TS*>(c.f0));
  p.bc = c.bc;
  p.miss = c.miss;
  p.table = c.table;
  p.nx = c.nx;
  p.ny = c.ny;
  p.nz = c.nz;
  p.x_begin = c.x_begin;
  p.plane = (long long)c.ny * c.nz;
  p.n = p.plane * c.nx;
  p.omega = c.omega;
  p.ghost_lo = static_cast<const TS*>(c.ghost_lo);
  p.ghost_hi = static_cast<const TS*>(c.ghost_hi);
  p.out_lo = static_cast<TS*>(c.out_lo);
  p.out_hi = static_cast<TS*>(c.out_hi);
  return launch_step<L, COLL, TC, TS>(p, c.x_count, c.requested_v, c.stream);
}

#define XLBN_DEFINE_STEP_DISPATCH(LAT, COLL)                                                                      \
  template <>                                                                                                     \
  int dispatch_step<LAT, COLL>(const StepCall& c) {                                                               \
    if (c.compute_dtype == XLBN_F32 && c.store_dtype == XLBN_F32) return run_step_typed<LAT, COLL, float, float>(c);   \
    if (c.compute_dtype == XLBN_F32 && c.store_dtype == XLBN_F16) return run_step_typed<LAT, COLL, float, __half>(c);  \
    if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F64) return run_step_typed<LAT, COLL, double, double>(c); \
    if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F32) return run_step_typed<LAT, COLL, double, float>(c);  \
    if (c.compute_dtype == XLBN_F64 && c.store_dtype == XLBN_F16) return run_step_typed<LAT, COLL, double, __half>(c); \
    return fail(XLBN_E_DTYPE, "unsupported precision policy: compute=%d store=%d", c.compute_dtype, c.store_dtype);    \
  }

}  // namespace xlbn
