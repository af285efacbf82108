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

#include <cstring>
#include <initializer_list>

#include "lbm_math.cuh"

namespace xlbn {

constexpr int kMaxQ = 27;

struct BcEntry {
  int kind;
  int pad;
  double rho;
  double u[3];
  // EquilibriumBC only: the cell's complete update, collide(feq(rho, u)), which does not depend on the pulled populations.
  // Refreshed by bc_precompute_kernel whenever omega changes; read by the half2-state pair path.
  float eq_out[kMaxQ];
  int pad2;
};

// Kernel parameters.  Everything that depends only on (population, x-plane class) is folded into pointer tables on the
// host, so that inside the kernel an address is ONE table entry (constant bank) + ONE 32-bit per-thread element offset
//   off = x * plane + y_src * nz + z_src
// (2 integer instructions per load/store instead of 64-bit multiplies per population).
//   pull[0][l] = f0 + l*n - ck0(l)*plane              planes whose x-neighbours are inside the array
//   pull[1][l] = plane x = 0   : c_x = +1 populations come from ghost_lo (or wrap to plane nx-1)
//   pull[2][l] = plane x = nx-1: c_x = -1 populations come from ghost_hi (or wrap to plane 0)
//   push[l]    = f1 + l*n
//   peer_hi[l] / peer_lo[l]: neighbour GPUs' ghost planes (biased by -x*plane), only for the face populations
template <class TS>
struct StepParams {
  const TS* pull[3][kMaxQ];
  TS* push[kMaxQ];
  TS* peer_hi[kMaxQ];  // written by plane nx-1 (c_x = +1 populations) or NULL
  TS* peer_lo[kMaxQ];  // written by plane 0    (c_x = -1 populations) or NULL
  const uint8_t* bc;
  uint8_t kinds[256];  // bc id -> xlbn_bc_kind (constant bank; 0 for ids without a BC), so common kinds are handled inline
  // boundary-cell path only
  const TS* f0;
  TS* f1;
  TS* f0w;  // writable alias of f0: aux recovery (nse_stepper.py:338)
  const uint32_t* miss;
  const BcEntry* table;  // 256 entries, indexed by bc id
  const TS* ghost_lo;    // plane "x = -1": populations with ck(0) = +1, [n_xdir][ny][nz]; NULL -> periodic wrap
  const TS* ghost_hi;    // plane "x = nx"
  int nx, ny, nz, x_begin;
  long long plane;  // ny * nz
  long long n;      // nx * ny * nz  (population stride)
  double omega;
  // extended collision models only (appended: the layout above is what the base kernels were validated with)
  double force[3];     // ForcedCollision / ExactDifference body force
  double smagorinsky;  // SmagorinskyLESBGK coefficient
  // EquilibriumBC at the INPUT side: feq(rho_bc, u_bc) of up to kEqSlots boundary ids, evaluated once on the host in the compute dtype
  // (same expression, same IEEE roundings as equilibrium<>()).  A thread that owns such a cell replaces the pulled populations by
  // these constants and takes the ordinary straight-line path — bc_equilibrium.py:76-86 followed by the collision, without a second
  // copy of the collision code executed by a whole warp for one lid lane.
  double eq_in[4][kMaxQ];
  uint8_t eq_ids[4];  // boundary id of each slot, 0 = unused
};
constexpr int kEqSlots = 4;

// ---- explicit global-space memory instructions (SASS: LDG / STG; the compiler cannot prove the address space of
//      pointers that arrive inside a parameter struct and would emit generic LD / ST) --------------------------------
template <int BYTES>
struct GMem;
template <>
struct GMem<1> {
  using T = uint8_t;
  static XLBN_DEV T ld(const void* p) {
#if XLBN_ON_HOST
    return *static_cast<const T*>(p);
#else
    uint32_t v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return (T)v;
#endif
  }
  static XLBN_DEV void st(void* p, T v) {
#if XLBN_ON_HOST
    *static_cast<T*>(p) = v;
#else
    asm volatile("st.global.u8 [%0], %1;" ::"l"(p), "r"((uint32_t)v) : "memory");
#endif
  }
};
template <>
struct GMem<2> {
  using T = uint16_t;
  static XLBN_DEV T ld(const void* p) {
#if XLBN_ON_HOST
    return *static_cast<const T*>(p);
#else
    T v;
    asm volatile("ld.global.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
#endif
  }
  static XLBN_DEV void st(void* p, T v) {
#if XLBN_ON_HOST
    *static_cast<T*>(p) = v;
#else
    asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"(v) : "memory");
#endif
  }
};
template <>
struct GMem<4> {
  using T = uint32_t;
  static XLBN_DEV T ld(const void* p) {
#if XLBN_ON_HOST
    return *static_cast<const T*>(p);
#else
    T v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
#endif
  }
  static XLBN_DEV void st(void* p, T v) {
#if XLBN_ON_HOST
    *static_cast<T*>(p) = v;
#elif XLBN_ST_CS
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    asm volatile("st.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
  }
};
template <>
struct GMem<8> {
  using T = uint2;
  static XLBN_DEV T ld(const void* p) {
#if XLBN_ON_HOST
    return *static_cast<const T*>(p);
#else
    T v;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
#endif
  }
  static XLBN_DEV void st(void* p, T v) {
#if XLBN_ON_HOST
    *static_cast<T*>(p) = v;
#else
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
#endif
  }
};
template <>
struct GMem<16> {
  using T = uint4;
  static XLBN_DEV T ld(const void* p) {
#if XLBN_ON_HOST
    return *static_cast<const T*>(p);
#else
    T v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#endif
  }
  static XLBN_DEV void st(void* p, T v) {
#if XLBN_ON_HOST
    *static_cast<T*>(p) = v;
#else
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
  }
};

template <class T, int V>
XLBN_DEV Pack<T, V> gload(const T* p) {
  using G = GMem<(int)sizeof(T) * V>;
  union {
    typename G::T raw;
    Pack<T, V> pack;
  } u;
  u.raw = G::ld(p);
  return u.pack;
}
template <class T, int V>
XLBN_DEV void gstore(T* p, const Pack<T, V>& x) {
  using G = GMem<(int)sizeof(T) * V>;
  union {
    typename G::T raw;
    Pack<T, V> pack;
  } u;
  u.pack = x;
  G::st(p, u.raw);
}

// FAST (reciprocal-based) divisions where the kernel is issue-bound: single-precision KBC.  fp64 and BGK keep IEEE division.
template <int COLL, class TC>
constexpr bool kFast = (kBaseCollision<COLL> == XLBN_KBC) && (sizeof(TC) == 4) && (COLL & kExactKbc) == 0;

// BGK / KBC take the two-argument form they were tuned and validated with; SmagorinskyLESBGK and every forced operator
// read their extra constants from the parameter block.
template <int COLL>
constexpr bool kExtCollision = COLL > XLBN_KBC;
template <class L, int COLL, class TC, class TS>
XLBN_DEV void collide_in_step(const StepParams<TS>& p, TC (&f)[L::Q], TC omega) {
  if constexpr (kExtCollision<COLL>) collide_cell_ext<L, COLL, TC, kFast<COLL, TC>>(f, omega, p.force, p.smagorinsky);
  else collide_cell<L, COLL, TC, kFast<COLL, TC>>(f, omega);
}

// f0[l] at an arbitrary (possibly out-of-range) kernel-coordinate cell: periodic in y/z; x through ghost or wrap.
template <class L, class TC, class TS>
XLBN_DEVFN TC load_f0_any(const StepParams<TS>& p, int l, int ck0, int slot, int x, int y, int z) {
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
XLBN_DEVFN __noinline__ void bc_cell(const StepParams<TS>& p, int id, int x, int y, int z, TC* fio) {
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
    XLBN_FOR(L::D, d) nk[L::kaxis(d)] = ni[d]; XLBN_END
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
    collide_in_step<L, COLL, TC, TS>(p, f, omega);
    XLBN_FOR(Q, l)
      if ((miss >> l) & 1u) f[L::opp(l)] = aux[l];
    XLBN_END
  } else {
    collide_in_step<L, COLL, TC, TS>(p, f, omega);
  }
  XLBN_FOR(Q, l) fio[l] = f[l]; XLBN_END
}

// ---- tuning knobs (compile-time; the defaults are the values selected on B200, see profiles/ and DESIGN.md) -------
#ifndef XLBN_MINB_OVERRIDE
#define XLBN_MINB_OVERRIDE 0  // > 0: force __launch_bounds__(128, N) for every variant
#endif
#ifndef XLBN_PREFETCH
#define XLBN_PREFETCH 0  // 1: prefetch.global.L2 the next x-plane's populations while this plane is processed
#endif
#ifndef XLBN_ST_CS
#define XLBN_ST_CS 0  // 1: st.global.cs (evict-first) for the population stores
#endif

template <class L, int COLL, class TC, class TS, int V, int MODE = 0>
struct StepTraits {
  static constexpr bool PK = MODE == 1;
  static constexpr int kThreads = 128;
  // Occupancy-first register budget (the kernel is latency-bound until ~48 warps/SM are resident, profiles/):
  // V*Q population registers (x2 for fp64) + collision temporaries + addresses.
  static constexpr int kW = (int)(sizeof(TC) / 4);
  static constexpr int kRegs = (COLL & kLeanKbc) ? (L::Q + 45) * kW  // lean KBC: f[] + (rho, u, usqr, pi / sv, sums) and addresses
                               : (MODE == 2 || MODE == 5) ? L::Q + 45  // populations stay packed as half2: q registers for two cells
                               : PK      ? V * L::Q + (kBaseCollision<COLL> == XLBN_KBC ? 4 : 2) * L::Q + 29  // pair temporaries take two registers each
                                         : V * L::Q * kW + (kBaseCollision<COLL> == XLBN_KBC ? L::Q * kW + 24 : 24) + (kForcedCollision<COLL> ? 8 * kW : 0) +
                                               (kBaseCollision<COLL> == XLBN_SMAGORINSKY_LES_BGK ? 8 * kW : 0) + 5;
  static constexpr int kMinBlocksRaw = 65536 / (kThreads * (kRegs > 255 ? 255 : kRegs));
  static constexpr int kMinBlocksAuto = kMinBlocksRaw < 1 ? 1 : (kMinBlocksRaw > 12 ? 12 : kMinBlocksRaw);
  // fp64 compute, D3Q19 BGK, one cell per thread: 8 CTAs/SM (64 registers) measured 3 % faster than the 7 the budget formula gives and
  // than 9 (profiles/r1_sweep_fp16_fp64_occupancy.txt)
  static constexpr bool kF64Q19 = sizeof(TC) == 8 && L::Q == 19 && V == 1 && MODE == 0 && COLL == XLBN_BGK;
  static constexpr int kMinBlocks = XLBN_MINB_OVERRIDE > 0 ? XLBN_MINB_OVERRIDE : (kF64Q19 ? 8 : kMinBlocksAuto);
};

// Store V cells (fused compute -> store conversion).  The outgoing face populations of planes 0 / nx-1 ALSO go straight
// into the neighbour GPUs' ghost planes through peer-mapped pointers.  MASKED: cells with id 255 are not written.
template <class L, class TC, class TS, int V, int XC, bool MASKED>
XLBN_DEV void store_cells(const StepParams<TS>& p, unsigned cell, const Pack<uint8_t, V>& ids, const TC (&f)[V][L::Q]) {
  if constexpr (!MASKED) {
    XLBN_FOR(L::Q, l)
      Pack<TS, V> a;
#pragma unroll
      for (int v = 0; v < V; ++v) a.v[v] = Cvt<TC, TS>::down(f[v][l]);
      gstore<TS, V>(p.push[l] + cell, a);
      if constexpr (L::ck(0, l) == 1 && (XC & 2)) {
        if (p.peer_hi[l]) gstore<TS, V>(p.peer_hi[l] + cell, a);
      } else if constexpr (L::ck(0, l) == -1 && (XC & 1)) {
        if (p.peer_lo[l]) gstore<TS, V>(p.peer_lo[l] + cell, a);
      }
    XLBN_END
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (ids.v[v] == 255) continue;
      XLBN_FOR(L::Q, l)
        Pack<TS, 1> a;
        a.v[0] = Cvt<TC, TS>::down(f[v][l]);
        gstore<TS, 1>(p.push[l] + (cell + v), a);
        if constexpr (L::ck(0, l) == 1 && (XC & 2)) {
          if (p.peer_hi[l]) gstore<TS, 1>(p.peer_hi[l] + (cell + v), a);
        } else if constexpr (L::ck(0, l) == -1 && (XC & 1)) {
          if (p.peer_lo[l]) gstore<TS, 1>(p.peer_lo[l] + (cell + v), a);
        }
      XLBN_END
    }
  }
}

// Boundary handling of a thread's V cells IN REGISTERS: f holds the pulled populations on entry and the values to store on exit
// (cells with id 255 are left as they are: the caller must not store them).
template <class L, int COLL, class TC, class TS, int V>
XLBN_DEV void bc_compute(const StepParams<TS>& p, const Pack<uint8_t, V>& ids, const int x, const int y, const int z0, TC (&f)[V][L::Q]) {
  constexpr int Q = L::Q;
  const TC omega = (TC)p.omega;
  // Threads with boundary cells.  The two kinds that make up closed-box walls and lids are handled in registers:
  //   FullwayBounceBack: out[l] = f_post_stream[opp[l]], no collision needed (bc_fullway_bounce_back.py:60-72)
  //   EquilibriumBC    : f = feq(rho_bc, u_bc), then the ordinary collision (bc_equilibrium.py:76-86)
  // every other kind goes through the out-of-line boundary-cell routine.
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int id = ids.v[v];
    if (id == 255) continue;
    int kind = id ? (int)p.kinds[id] : 0;
    if (kind == XLBN_BC_EQUILIBRIUM) {
      const BcEntry* e = p.table + id;
      TC u[L::D];
      XLBN_FOR(L::D, d) u[d] = (TC)e->u[d]; XLBN_END
      equilibrium<L, TC>((TC)e->rho, u, f[v]);
      kind = XLBN_BC_NONE;
    }
    if (kind == XLBN_BC_NONE) {
      collide_in_step<L, COLL, TC, TS>(p, f[v], omega);
    } else if (kind == XLBN_BC_FULLWAY_BOUNCE_BACK) {
      TC t[Q];
      XLBN_FOR(Q, l) t[l] = f[v][L::opp(l)]; XLBN_END
      XLBN_FOR(Q, l) f[v][l] = t[l]; XLBN_END
    } else {
      TC tmp[Q];
      XLBN_FOR(Q, l) tmp[l] = f[v][l]; XLBN_END
      bc_cell<L, COLL, TC, TS>(p, id, x, y, z0 + v, tmp);
      XLBN_FOR(Q, l) f[v][l] = tmp[l]; XLBN_END
    }
  }
}

template <class L, int COLL, class TC, class TS, int V, int XC>
XLBN_DEV void bc_tail(const StepParams<TS>& p, const Pack<uint8_t, V>& ids, const int x, const int y, const int z0, const unsigned cell,
                      const bool any_solid, TC (&f)[V][L::Q]) {
  bc_compute<L, COLL, TC, TS, V>(p, ids, x, y, z0, f);
  if (any_solid) store_cells<L, TC, TS, V, XC, true>(p, cell, ids, f);
  else store_cells<L, TC, TS, V, XC, false>(p, cell, ids, f);
}

// XC = x-plane class of this block: 0 interior, 1 plane 0, 2 plane nx-1, 3 both (nx == 1)
template <class L, int COLL, class TC, class TS, int V, int XC>
XLBN_DEV void step_body(const StepParams<TS>& p, const int x, const int y, const int z0) {
  constexpr int Q = L::Q;
  const unsigned nz = (unsigned)p.nz;
  const unsigned xoff = (unsigned)x * (unsigned)p.plane;
  // rows the pull reads from: y - cy with periodic wrap (stream.py:66-78); element offsets inside the population
  const unsigned row_c = xoff + (unsigned)y * nz;
  const unsigned row_m = xoff + (unsigned)(y == 0 ? p.ny - 1 : y - 1) * nz;  // source row for cy = +1
  const unsigned row_p = xoff + (unsigned)(y == p.ny - 1 ? 0 : y + 1) * nz;  // source row for cy = -1
  const unsigned z_lo = (z0 == 0) ? nz - 1 : (unsigned)z0 - 1;                 // element left of the vector  (cz = +1)
  const unsigned z_hi = ((unsigned)z0 + V >= nz) ? 0u : (unsigned)z0 + V;      // element right of the vector (cz = -1)
  const unsigned cell = row_c + (unsigned)z0;

  // boundary ids of the V cells (one V-byte load)
  const Pack<uint8_t, V> ids = gload<uint8_t, V>(p.bc + cell);
  bool any_solid = false, all_solid = true, any_bc = false;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    any_solid |= (ids.v[v] == 255);
    all_solid &= (ids.v[v] == 255);
    any_bc |= (ids.v[v] != 0);
  }
  if (all_solid) return;  // nse_stepper.py:356-358

  TC f[V][Q];
  XLBN_FOR(Q, l)
    constexpr int cx = L::ck(0, l), cy = L::ck(1, l), cz = L::ck(2, l);
    constexpr int tab = (cx == 1 && (XC & 1)) ? 1 : ((cx == -1 && (XC & 2)) ? 2 : 0);
    const TS* base = p.pull[tab][l];
    const unsigned row = (cy == 1 ? row_m : (cy == -1 ? row_p : row_c));
    if constexpr (cz == 0) {
      const Pack<TS, V> a = gload<TS, V>(base + (row + (unsigned)z0));
#if XLBN_PREFETCH
      if constexpr (XC == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (row + (unsigned)z0) + p.plane));
#endif
#pragma unroll
      for (int v = 0; v < V; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v]);
    } else if constexpr (V == 1) {
      const Pack<TS, 1> a = gload<TS, 1>(base + (row + (cz == 1 ? z_lo : z_hi)));
#if XLBN_PREFETCH
      if constexpr (XC == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (row + (cz == 1 ? z_lo : z_hi)) + p.plane));
#endif
      f[0][l] = Cvt<TC, TS>::up(a.v[0]);
    } else if constexpr (cz == 1) {  // out[z] = in[z - 1]
      const Pack<TS, V> a = gload<TS, V>(base + (row + (unsigned)z0));
      const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_lo));
      f[0][l] = Cvt<TC, TS>::up(e.v[0]);
#pragma unroll
      for (int v = 1; v < V; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v - 1]);
    } else {  // out[z] = in[z + 1]
      const Pack<TS, V> a = gload<TS, V>(base + (row + (unsigned)z0));
      const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_hi));
#pragma unroll
      for (int v = 0; v < V - 1; ++v) f[v][l] = Cvt<TC, TS>::up(a.v[v + 1]);
      f[V - 1][l] = Cvt<TC, TS>::up(e.v[0]);
    }
  XLBN_END

  const TC omega = (TC)p.omega;
  if (any_bc) {  // EquilibriumBC cells whose feq is in the parameter block: f = feq(rho_bc, u_bc), then they are ordinary cells
    bool rest = false;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int id = ids.v[v];
      if (id == 0) continue;
      int slot = -1;
#pragma unroll
      for (int i = 0; i < kEqSlots; ++i)
        if (p.eq_ids[i] == id) slot = i;
      if (slot < 0) {
        rest = true;
        continue;
      }
      XLBN_FOR(Q, l) f[v][l] = (TC)p.eq_in[slot][l]; XLBN_END
    }
    any_bc = rest;
  }
  if (!any_bc) {
    // straight-line path: no boundary cell among this thread's V cells (ends here, so that its register allocation is
    // independent of the boundary code)
#pragma unroll
    for (int v = 0; v < V; ++v) collide_in_step<L, COLL, TC, TS>(p, f[v], omega);
    store_cells<L, TC, TS, V, XC, false>(p, cell, ids, f);
    return;
  }
  bc_tail<L, COLL, TC, TS, V, XC>(p, ids, x, y, z0, cell, any_solid, f);
}

// ---- packed pair path: two neighbouring cells per fp32x2 register pair (FADD2 / FMUL2 / FFMA2) ---------------------
template <class TS>
XLBN_DEV f32x2 pair_up(TS lo, TS hi);
template <>
XLBN_DEV f32x2 pair_up<float>(float lo, float hi) { return f32x2(lo, hi); }
template <>
XLBN_DEV f32x2 pair_up<__half>(__half lo, __half hi) { return f32x2(__half22float2(__halves2half2(lo, hi))); }
template <class TS>
XLBN_DEV void pair_down(f32x2 v, TS& lo, TS& hi);
template <>
XLBN_DEV void pair_down<float>(f32x2 v, float& lo, float& hi) { lo = v.v.x; hi = v.v.y; }
template <>
XLBN_DEV void pair_down<__half>(f32x2 v, __half& lo, __half& hi) {
  const __half2 h = __float22half2_rn(v.v);
  lo = __low2half(h);
  hi = __high2half(h);
}

template <class L, int COLL, class TS, int V, int XC>
XLBN_DEV void step_body_pk(const StepParams<TS>& p, const int x, const int y, const int z0) {
  static_assert(V % 2 == 0, "pair path needs an even number of cells per thread");
  constexpr int Q = L::Q, NP = V / 2;
  const unsigned nz = (unsigned)p.nz;
  const unsigned xoff = (unsigned)x * (unsigned)p.plane;
  const unsigned row_c = xoff + (unsigned)y * nz;
  const unsigned row_m = xoff + (unsigned)(y == 0 ? p.ny - 1 : y - 1) * nz;
  const unsigned row_p = xoff + (unsigned)(y == p.ny - 1 ? 0 : y + 1) * nz;
  const unsigned z_lo = (z0 == 0) ? nz - 1 : (unsigned)z0 - 1;
  const unsigned z_hi = ((unsigned)z0 + V >= nz) ? 0u : (unsigned)z0 + V;
  const unsigned cell = row_c + (unsigned)z0;

  const Pack<uint8_t, V> ids = gload<uint8_t, V>(p.bc + cell);
  bool any_solid = false, all_solid = true, any_bc = false;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    any_solid |= (ids.v[v] == 255);
    all_solid &= (ids.v[v] == 255);
    any_bc |= (ids.v[v] != 0);
  }
  if (all_solid) return;

  f32x2 f[NP][Q];
  XLBN_FOR(Q, l)
    constexpr int cx = L::ck(0, l), cy = L::ck(1, l), cz = L::ck(2, l);
    constexpr int tab = (cx == 1 && (XC & 1)) ? 1 : ((cx == -1 && (XC & 2)) ? 2 : 0);
    const TS* base = p.pull[tab][l];
    const unsigned row = (cy == 1 ? row_m : (cy == -1 ? row_p : row_c));
    const Pack<TS, V> a = gload<TS, V>(base + (row + (unsigned)z0));
    if constexpr (cz == 0) {
#pragma unroll
      for (int j = 0; j < NP; ++j) f[j][l] = pair_up<TS>(a.v[2 * j], a.v[2 * j + 1]);
    } else if constexpr (cz == 1) {  // out[z] = in[z - 1]
      const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_lo));
      f[0][l] = pair_up<TS>(e.v[0], a.v[0]);
#pragma unroll
      for (int j = 1; j < NP; ++j) f[j][l] = pair_up<TS>(a.v[2 * j - 1], a.v[2 * j]);
    } else {  // out[z] = in[z + 1]
      const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_hi));
#pragma unroll
      for (int j = 0; j < NP - 1; ++j) f[j][l] = pair_up<TS>(a.v[2 * j + 1], a.v[2 * j + 2]);
      f[NP - 1][l] = pair_up<TS>(a.v[V - 1], e.v[0]);
    }
  XLBN_END

  if (!any_bc) {
    const f32x2 omega((float)p.omega);
#pragma unroll
    for (int j = 0; j < NP; ++j) collide_cell<L, COLL, f32x2, true>(f[j], omega);
    XLBN_FOR(Q, l)
      Pack<TS, V> a;
#pragma unroll
      for (int j = 0; j < NP; ++j) pair_down<TS>(f[j][l], a.v[2 * j], a.v[2 * j + 1]);
      gstore<TS, V>(p.push[l] + cell, a);
      if constexpr (L::ck(0, l) == 1 && (XC & 2)) {
        if (p.peer_hi[l]) gstore<TS, V>(p.peer_hi[l] + cell, a);
      } else if constexpr (L::ck(0, l) == -1 && (XC & 1)) {
        if (p.peer_lo[l]) gstore<TS, V>(p.peer_lo[l] + cell, a);
      }
    XLBN_END
    return;
  }
  // threads with boundary cells: unpack and take the scalar boundary tail
  float fs[V][Q];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    XLBN_FOR(Q, l)
      fs[2 * j][l] = f[j][l].v.x;
      fs[2 * j + 1][l] = f[j][l].v.y;
    XLBN_END
  }
  bc_tail<L, COLL, float, TS, V, XC>(p, ids, x, y, z0, cell, any_solid, fs);
}


// One thread per bc id: EquilibriumBC entries get their constant cell update (same device functions as every other path).
template <class L, int COLL>
__global__ void bc_precompute_kernel(BcEntry* table, float omega) {
  const int id = threadIdx.x;
  if (id >= 256 || table[id].kind != XLBN_BC_EQUILIBRIUM) return;
  float u[L::D], f[L::Q];
  XLBN_FOR(L::D, d) u[d] = (float)table[id].u[d]; XLBN_END
  equilibrium<L, float>((float)table[id].rho, u, f);
  collide_cell<L, COLL, float, kFast<COLL, float>>(f, omega);
  XLBN_FOR(L::Q, l) table[id].eq_out[l] = f[l]; XLBN_END
}

// The arithmetic of the half2-state paths: moments -> equilibrium -> BGK for the two cells held in h (packed fp32x2: both cells at once),
// emit(l, out) receives the relaxed pair of population l.  One IEEE rounding per operation, in the reference's order (csrc/lbm_math.cuh
// "ROUNDINGS"): bit-identical to the scalar path and to the reference kernel.
template <class L, class Emit>
XLBN_DEV void h2_collide_each(const __half2 (&h)[L::Q], const float omega_f, Emit&& emit) {
  constexpr int Q = L::Q;
  // moments (macroscopic.py:43-47)
  f32x2 rho(0.0f), u[L::D];
  XLBN_FOR(L::D, d) u[d] = f32x2(0.0f); XLBN_END
  XLBN_FOR(Q, l)
    const f32x2 f(__half22float2(h[l]));
    rho += f;
    XLBN_FOR(L::D, d)
      if constexpr (L::c(d, l) == 1) u[d] += f;
      else if constexpr (L::c(d, l) == -1) u[d] -= f;
    XLBN_END
  XLBN_END
  // u = (sum c f) / rho, correctly rounded (first_moment.py:38)
  const f32x2 inv = rcp_refined_(rho);
  XLBN_FOR(L::D, d) u[d] = div_by_(u[d], rho, inv); XLBN_END
  // mul_then_add_: a product that is rounded BEFORE the addition it feeds (see common.cuh: the compiler would contract the pair)
  f32x2 uu = mul_then_add_(u[0], u[0]);
  XLBN_FOR(L::D - 1, d) uu = uu + mul_then_add_(u[d + 1], u[d + 1]); XLBN_END
  const f32x2 usqr = mul_then_add_(f32x2(1.5f), uu);
  const f32x2 omega(omega_f);
  // equilibrium + BGK relaxation per population (quadratic_equilibrium.py:35-60, bgk.py:30-34).  Opposite directions are relaxed
  // together: 3 c_opp.u = -(3 c.u) exactly (IEEE addition and multiplication are sign-symmetric), so the pair shares its cu.
  auto relax = [&](auto l_, const f32x2 cu) {
    constexpr int l = decltype(l_)::value;
    const f32x2 f(__half22float2(h[l]));
    // 1 + 0.5 cu may contract: the product is exact
    const f32x2 feq = mul_then_add_(rho * f32x2(L::w(l)), f32x2(1.0f) + mul_then_add_(cu, f32x2(1.0f) + f32x2(0.5f) * cu) - usqr);
    emit(l_, f - mul_then_add_(omega, f - feq));
  };
  XLBN_FOR(Q, l)
    if constexpr (L::opp(l) >= l) {
      f32x2 cu(0.0f);
      XLBN_FOR(L::D, d)
        if constexpr (L::c(d, l) == 1) cu += u[d];
        else if constexpr (L::c(d, l) == -1) cu -= u[d];
      XLBN_END
      cu *= f32x2(3.0f);
      relax(l_, cu);
      if constexpr (L::opp(l) != l) relax(IC<L::opp(l)>{}, -cu);
    }
  XLBN_END
}

// moments + equilibrium + BGK + narrow + store for the two cells of a half2-state thread.  BCV = 2: per-half handling of
// FullwayBounceBack (bit copy of the opposite population's half; fp16 -> fp32 -> fp16 is exact) and EquilibriumBC cells
// (the precomputed constant update, BcEntry::eq_out); BCV = 1: FullwayBounceBack only (no table pointers, no constant
// loads: the walls of a closed box); BCV = 0: no boundary cell.  id_lo / id_hi = bc ids of the two cells.
template <class L, int XC, int BCV>
XLBN_DEV void h2_collide_store(const StepParams<__half>& p, const __half2 (&h)[L::Q], const unsigned cell, const int id_lo, const int id_hi) {
  using TS = __half;
  bool eq_lo = false, eq_hi = false, fw_lo = false, fw_hi = false;
  const float* out_lo = nullptr;
  const float* out_hi = nullptr;
  if constexpr (BCV == 2) {
    const int k_lo = id_lo ? (int)p.kinds[id_lo] : 0, k_hi = id_hi ? (int)p.kinds[id_hi] : 0;
    eq_lo = k_lo == XLBN_BC_EQUILIBRIUM;
    eq_hi = k_hi == XLBN_BC_EQUILIBRIUM;
    fw_lo = k_lo == XLBN_BC_FULLWAY_BOUNCE_BACK;
    fw_hi = k_hi == XLBN_BC_FULLWAY_BOUNCE_BACK;
    out_lo = p.table[id_lo].eq_out;
    out_hi = p.table[id_hi].eq_out;
  } else if constexpr (BCV == 1) {  // the warp holds fluid and FullwayBounceBack cells only
    fw_lo = id_lo && p.kinds[id_lo] == XLBN_BC_FULLWAY_BOUNCE_BACK;
    fw_hi = id_hi && p.kinds[id_hi] == XLBN_BC_FULLWAY_BOUNCE_BACK;
  }
  h2_collide_each<L>(h, (float)p.omega, [&](auto l_, f32x2 out) {
    constexpr int l = decltype(l_)::value;
    if constexpr (BCV == 2) {  // bc_equilibrium.py:76-86 followed by the ordinary collision = a per-BC constant
      if (eq_lo) out.v.x = out_lo[l];
      if (eq_hi) out.v.y = out_hi[l];
    }
    __half2 o = __float22half2_rn(out.v);
    if constexpr (BCV != 0) {  // bc_fullway_bounce_back.py:60-72: out[l] = f_post_stream[opp l]
      if (fw_lo) o = __halves2half2(__low2half(h[L::opp(l)]), __high2half(o));
      if (fw_hi) o = __halves2half2(__low2half(o), __high2half(h[L::opp(l)]));
    }
    Pack<TS, 2> a;
    a.v[0] = __low2half(o);
    a.v[1] = __high2half(o);
    gstore<TS, 2>(p.push[l] + cell, a);
    if constexpr (L::ck(0, l) == 1 && (XC & 2)) {
      if (p.peer_hi[l]) gstore<TS, 2>(p.peer_hi[l] + cell, a);
    } else if constexpr (L::ck(0, l) == -1 && (XC & 1)) {
      if (p.peer_lo[l]) gstore<TS, 2>(p.peer_lo[l] + cell, a);
    }
  });
}

// ---- half2-state pair path (FP32FP16, BGK): two cells per thread, populations kept as the loaded half2 words ---------------
// The q post-stream populations of the two cells stay in q 32-bit registers (half2) — the storage format IS the register
// format — and are widened to a fp32x2 pair on the fly twice: once to accumulate rho and u, once for equilibrium +
// relaxation, whose result is narrowed (one F2FP) and stored immediately.  Live state: q + ~25 registers for two cells,
// so the kernel keeps the residency of the one-cell path while issuing ~2.5x fewer instructions per cell (FADD2 / FMUL2 /
// FFMA2 for both cells at once); the fp16 path is issue-bound otherwise (profiles/README.md).
// SPLIT (tuning variant, cells_per_thread = 203): warps whose boundary cells are all FullwayBounceBack take a leaner boundary
// variant without the EquilibriumBC table pointers / constant loads.
template <class L, int XC, bool SPLIT = false>
XLBN_DEV void step_body_h2(const StepParams<__half>& p, const int x, const int y, const int z0) {
  using TS = __half;
  constexpr int Q = L::Q, V = 2;
  const unsigned nz = (unsigned)p.nz;
  const unsigned xoff = (unsigned)x * (unsigned)p.plane;
  const unsigned row_c = xoff + (unsigned)y * nz;
  const unsigned row_m = xoff + (unsigned)(y == 0 ? p.ny - 1 : y - 1) * nz;
  const unsigned row_p = xoff + (unsigned)(y == p.ny - 1 ? 0 : y + 1) * nz;
  const unsigned z_lo = (z0 == 0) ? nz - 1 : (unsigned)z0 - 1;
  const unsigned z_hi = ((unsigned)z0 + V >= nz) ? 0u : (unsigned)z0 + V;
  const unsigned cell = row_c + (unsigned)z0;

  const Pack<uint8_t, V> ids = gload<uint8_t, V>(p.bc + cell);
  const bool any_solid = (ids.v[0] == 255) | (ids.v[1] == 255);
  const bool any_bc = (ids.v[0] != 0) | (ids.v[1] != 0);
  if ((ids.v[0] == 255) & (ids.v[1] == 255)) return;

  // Three separate code paths, each with its own load phase, so that the register allocation of the straight-line path
  // is independent of the boundary code (the branch is taken BEFORE anything is loaded).
  auto load_all = [&](__half2 (&h)[Q]) {
    XLBN_FOR(Q, l)
      constexpr int cx = L::ck(0, l), cy = L::ck(1, l), cz = L::ck(2, l);
      constexpr int tab = (cx == 1 && (XC & 1)) ? 1 : ((cx == -1 && (XC & 2)) ? 2 : 0);
      const TS* base = p.pull[tab][l];
      const unsigned row = (cy == 1 ? row_m : (cy == -1 ? row_p : row_c));
      const Pack<TS, 2> a = gload<TS, 2>(base + (row + (unsigned)z0));
      if constexpr (cz == 0) {
        h[l] = __halves2half2(a.v[0], a.v[1]);
      } else if constexpr (cz == 1) {  // out[z] = in[z - 1]
        const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_lo));
        h[l] = __halves2half2(e.v[0], a.v[0]);
      } else {  // out[z] = in[z + 1]
        const Pack<TS, 1> e = gload<TS, 1>(base + (row + z_hi));
        h[l] = __halves2half2(a.v[1], e.v[0]);
      }
    XLBN_END
  };

  // Warp-uniform choice of the code path (threads of a warp never serialise through two paths):
  //   no boundary cell in the warp                      -> straight pair path
  //   only fluid / FullwayBounceBack / EquilibriumBC    -> pair path with per-half boundary handling
  //   anything else (other BC kinds, solid cells)       -> per-thread: pair path or scalar boundary tail
  const int k0 = ids.v[0] ? (int)p.kinds[ids.v[0]] : 0, k1 = ids.v[1] ? (int)p.kinds[ids.v[1]] : 0;
  const auto simple = [](int id, int k) { return id != 255 && (k == XLBN_BC_NONE || k == XLBN_BC_FULLWAY_BOUNCE_BACK || k == XLBN_BC_EQUILIBRIUM); };
  const unsigned active = XLBN_ACTIVEMASK();
  const bool warp_any_bc = XLBN_ANY(active, any_bc);
  if (!warp_any_bc) {
    __half2 h[Q];
    load_all(h);
    h2_collide_store<L, XC, 0>(p, h, cell, 0, 0);
    return;
  }
  if (XLBN_ALL(active, simple(ids.v[0], k0) && simple(ids.v[1], k1))) {
    if constexpr (SPLIT) {
      if (XLBN_ALL(active, k0 != XLBN_BC_EQUILIBRIUM && k1 != XLBN_BC_EQUILIBRIUM)) {
        __half2 h[Q];
        load_all(h);
        h2_collide_store<L, XC, 1>(p, h, cell, ids.v[0], ids.v[1]);
        return;
      }
    }
    __half2 h[Q];
    load_all(h);
    h2_collide_store<L, XC, 2>(p, h, cell, ids.v[0], ids.v[1]);
    return;
  }
  if (!any_bc) {
    __half2 h[Q];
    load_all(h);
    h2_collide_store<L, XC, 0>(p, h, cell, 0, 0);
    return;
  }
  __half2 h[Q];
  load_all(h);
  // threads with boundary cells: widen and take the scalar boundary tail
  float fs[V][Q];
  XLBN_FOR(Q, l)
    const float2 f = __half22float2(h[l]);
    fs[0][l] = f.x;
    fs[1][l] = f.y;
  XLBN_END
  bc_tail<L, XLBN_BGK, float, TS, V, XC>(p, ids, x, y, z0, cell, any_solid, fs);
}

template <class L, int COLL, class TC, class TS, int V, int MODE>
__global__ void __launch_bounds__(StepTraits<L, COLL, TC, TS, V, MODE>::kThreads, StepTraits<L, COLL, TC, TS, V, MODE>::kMinBlocks)
    step_kernel(const __grid_constant__ StepParams<TS> p) {
  const int zv = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = p.x_begin + blockIdx.z;
  const int z0 = zv * V;
  if (z0 >= p.nz || y >= p.ny) return;
  // block-uniform dispatch on the x-plane class: interior planes carry no ghost / wrap logic at all
  const bool first = (x == 0), last = (x == p.nx - 1);
  if constexpr (MODE == 2) {
    if (!first && !last) step_body_h2<L, 0>(p, x, y, z0);
    else if (first && !last) step_body_h2<L, 1>(p, x, y, z0);
    else if (last && !first) step_body_h2<L, 2>(p, x, y, z0);
    else step_body_h2<L, 3>(p, x, y, z0);
  } else if constexpr (MODE == 5) {
    if (!first && !last) step_body_h2<L, 0, true>(p, x, y, z0);
    else if (first && !last) step_body_h2<L, 1, true>(p, x, y, z0);
    else if (last && !first) step_body_h2<L, 2, true>(p, x, y, z0);
    else step_body_h2<L, 3, true>(p, x, y, z0);
  } else if constexpr (MODE == 1) {
    if (!first && !last) step_body_pk<L, COLL, TS, V, 0>(p, x, y, z0);
    else if (first && !last) step_body_pk<L, COLL, TS, V, 1>(p, x, y, z0);
    else if (last && !first) step_body_pk<L, COLL, TS, V, 2>(p, x, y, z0);
    else step_body_pk<L, COLL, TS, V, 3>(p, x, y, z0);
  } else {
    if (!first && !last) step_body<L, COLL, TC, TS, V, 0>(p, x, y, z0);
    else if (first && !last) step_body<L, COLL, TC, TS, V, 1>(p, x, y, z0);
    else if (last && !first) step_body<L, COLL, TC, TS, V, 2>(p, x, y, z0);
    else step_body<L, COLL, TC, TS, V, 3>(p, x, y, z0);
  }
}

}  // namespace xlbn
#include "step_tile.cuh"  // the persistent TMA-fed tile kernel (FP32FP16 BGK), built on the code above
namespace xlbn {

// ---- host-side launch ------------------------------------------------------------------------------------------------
template <class L, int COLL, class TC, class TS, int V, int MODE = 0>
int launch_step_v(const StepParams<TS>& p, int x_count, cudaStream_t stream) {
  constexpr int T = StepTraits<L, COLL, TC, TS, V, MODE>::kThreads;
  const int nzv = (p.nz + V - 1) / V;
  int bx = 32;
  while (bx < nzv && bx < T) bx *= 2;
  const int by = T / bx;
  dim3 block(bx, by, 1);
  dim3 grid((nzv + bx - 1) / bx, (p.ny + by - 1) / by, x_count);
  if (grid.y > 65535u || grid.z > 65535u) return fail(XLBN_E_SHAPE, "grid too large for launch: ny=%d x_count=%d", p.ny, x_count);
  step_kernel<L, COLL, TC, TS, V, MODE><<<grid, block, 0, stream>>>(p);
  XLBN_LAUNCH_OK("step_kernel launch");
  return 0;
}

// V must divide nz and keep every vector access aligned; otherwise fall back to the next smaller V.
inline int pick_cells_per_thread(int requested, int dflt, int esize, int nz, const void* bc, std::initializer_list<const void*> arrays) {
  int v = requested > 0 ? requested : dflt;
  const int vmax = 16 / esize;
  if (v > vmax) v = vmax;
  while (v & (v - 1)) --v;  // power of two
  auto aligned = [&](const void* q, size_t a) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) % a) == 0; };
  while (v > 1) {
    bool ok = (nz % v == 0) && aligned(bc, v);
    for (const void* q : arrays) ok = ok && aligned(q, (size_t)esize * v);
    if (ok) break;
    v /= 2;
  }
  return v;
}

// requested_v: 0 = library default; 1, 2, 4, 8 = scalar path with that many cells per thread; 102, 104 = packed pair
// path (fp32 compute only); 202 = half2-state pair path (FP32FP16 BGK only: the two cells' populations stay packed as
// half2 registers and are converted on the fly, once for the moments and once for the relaxation).
template <class L, int COLL, class TC, class TS>
int launch_step_base(const StepParams<TS>& p, int x_count, int requested_v, const void* f0, const void* f1, const void* g0, const void* g1,
                     const void* o0, const void* o1, BcEntry* table_rw, double* eq_omega_state, cudaStream_t stream) {
  constexpr bool can_pack = sizeof(TC) == 4 && sizeof(TS) <= 4;
  // defaults selected on B200 (profiles/)
  int req = requested_v;
  constexpr bool can_h2 = sizeof(TC) == 4 && sizeof(TS) == 2 && COLL == XLBN_BGK;
  if (req == 0) req = can_h2 ? 202 : 1;
#if !XLBN_ON_HOST
  // FP32FP16 BGK on 3-D lattices: the tile kernel (step_tile.cuh) wherever the slab can be tiled — 1024-cell tiles (16 consumer warps, one
  // CTA per SM) if they fit the plane, else 512-cell tiles (8 consumer warps, two CTAs per SM).  B200, 512^3 cavity: 0.93 / 0.87 of the
  // roofline against 0.70 for the direct-load pair path.  402 = that choice, explicitly; 404 = 512-cell tiles; 403 = 512-cell tiles, three
  // CTAs per SM (D3Q19; a tuning variant that spills).
  if constexpr (can_h2 && L::D == 3) {
    const bool peers = o0 != nullptr || o1 != nullptr;
    const bool fits1024 = tile_eligible<L, 1024>(p, f0, f1, g0, g1, peers), fits512 = tile_eligible<L, 512>(p, f0, f1, g0, g1, peers);
    if (requested_v == 0 && (fits1024 || fits512)) req = 402;
    if (req == 402 || req == 403 || req == 404) {
      if (!(req == 402 ? (fits1024 || fits512) : fits512))
        return fail(XLBN_E_SHAPE, "cells_per_thread = %d: the tile kernel needs nz | 512, nz %% 8 == 0, ny %% (512 / nz) == 0, 16-byte aligned arrays and no halo handle (nz = %d, ny = %d)", req, p.nz, p.ny);
      if (eq_omega_state && !(p.omega == *eq_omega_state)) {
        bc_precompute_kernel<L, COLL><<<1, 256, 0, stream>>>(table_rw, (float)p.omega);
        XLBN_LAUNCH_OK("bc_precompute_kernel");
        *eq_omega_state = p.omega;
      }
      if constexpr (L::Q <= 19) {
        if (req == 403) return launch_step_tile<L, 512, 3>(p, x_count, stream);
      }
      if (req == 402 && fits1024) return launch_step_tile<L, 1024, 1>(p, x_count, stream);
      return launch_step_tile<L, 512, 2>(p, x_count, stream);
    }
  } else {
    if (req == 402 || req == 403 || req == 404) return fail(XLBN_E_ARG, "cells_per_thread = %d: the tile kernel exists for FP32FP16 BGK on 3-D lattices only", req);
  }
#endif
#if !XLBN_ON_HOST
  // 501 / 502: the scalar tile kernel (step_tile.cuh: one cell per consumer thread, TMA-fed, persistent) with one / two CTAs per SM.
  // Built for the BGK operators with fp32 storage (FP32FP32, FP64FP32) on the 3-D lattices.
  // It is the default for D3Q19 BGK with fp32 storage wherever the slab can be tiled: FP32FP32 512^3 cavity 1.03 of the measured copy
  // bandwidth against 0.98 for the direct-load kernel, 256^3 1.00 against 0.96, 128^3 0.88 against 0.87; FP64FP32 0.74 against 0.71
  // (profiles/r2_call13_matrix.txt, r2_call15_matrix.txt).  D3Q27 (issue-bound with 16 warps: 0.87 against 1.00) keeps the direct kernel.
  if constexpr (COLL == XLBN_BGK && sizeof(TS) == 4 && L::D == 3 && L::Q == 19) {
    if (requested_v == 0 && tile1_eligible<L, TS>(p, f0, f1, g0, g1, o0 != nullptr || o1 != nullptr)) req = 501;
  }
  if (req == 501 || req == 502) {
    if constexpr (COLL == XLBN_BGK && sizeof(TS) == 4 && L::D == 3) {
      if (!tile1_eligible<L, TS>(p, f0, f1, g0, g1, o0 != nullptr || o1 != nullptr))
        return fail(XLBN_E_SHAPE, "cells_per_thread = %d: the scalar tile kernel needs nz | 512, nz %% 16 == 0, ny %% (512 / nz) == 0, 16-byte aligned arrays and no halo handle (nz = %d, ny = %d)", req, p.nz, p.ny);
      if constexpr (sizeof(TC) == 4) {
        if (req == 502) return launch_step_tile1<L, COLL, TC, TS, 2>(p, x_count, stream);
      }
      return launch_step_tile1<L, COLL, TC, TS, 1>(p, x_count, stream);
    } else {
      return fail(XLBN_E_ARG, "cells_per_thread = %d: the scalar tile kernel is built for BGK with fp32 storage on 3-D lattices", req);
    }
  }
#endif
  if (req == 202 || req == 203) {
    if constexpr (can_h2) {
      if (pick_cells_per_thread(2, 1, (int)sizeof(TS), p.nz, p.bc, {f0, f1, g0, g1, o0, o1}) == 2) {
        if (eq_omega_state && !(p.omega == *eq_omega_state)) {
          // first step / omega changed (it is a per-call argument and may ramp): recompute the EquilibriumBC constants.
          // Stream-ordered before the step launch below — no host synchronisation, capturable in a CUDA graph.  A caller that
          // launches ONE step on several streams (the slab path) publishes omega first with xlbn_stepper_prepare.
          bc_precompute_kernel<L, COLL><<<1, 256, 0, stream>>>(table_rw, (float)p.omega);
          XLBN_LAUNCH_OK("bc_precompute_kernel");
          *eq_omega_state = p.omega;
        }
        if (req == 203) return launch_step_v<L, COLL, TC, TS, 2, 5>(p, x_count, stream);
        return launch_step_v<L, COLL, TC, TS, 2, 2>(p, x_count, stream);
      }
      req = 1;  // odd nz or misaligned arrays: scalar fallback
    } else {
      return fail(XLBN_E_ARG, "cells_per_thread = 202: the half2-state path exists for FP32FP16 BGK only");
    }
  }  // one cell per thread at maximum residency won every comparison on B200 (profiles/r1_bench_matrix2.txt)
  bool packed = req >= 100;
  if (packed && !can_pack) return fail(XLBN_E_ARG, "cells_per_thread = %d: the packed pair path needs fp32 compute and fp32/fp16 storage", req);
  int v = pick_cells_per_thread(packed ? req - 100 : req, 1, (int)sizeof(TS), p.nz, p.bc, {f0, f1, g0, g1, o0, o1});
  if (packed && v < 2) packed = false;  // nz odd or misaligned: scalar fallback
  if constexpr (can_pack) {
    if (packed) return launch_step_v<L, COLL, TC, TS, 2, 1>(p, x_count, stream);  // wider pair variants spill (profiles/)
  }
  switch (v) {
    case 1: return launch_step_v<L, COLL, TC, TS, 1>(p, x_count, stream);
    case 2: return launch_step_v<L, COLL, TC, TS, 2>(p, x_count, stream);
    default:  // 4 and 8 (8 is served by the 4-wide variant; it never won on B200)
      if constexpr (sizeof(TS) <= 4) return launch_step_v<L, COLL, TC, TS, 4>(p, x_count, stream);
      else return launch_step_v<L, COLL, TC, TS, 2>(p, x_count, stream);
  }
}

// Extended collision models (SmagorinskyLESBGK, forced operators): one cell per thread — D3Q19 with fp32 storage can take the scalar
// tile kernel where the slab can be tiled (cells_per_thread 501; 0 picks it for unforced SmagorinskyLESBGK), everything else the
// direct-load kernel; other cells_per_thread values are ignored.
template <class L, int COLL, class TC, class TS>
int launch_step(const StepParams<TS>& p, int x_count, int requested_v, const void* f0, const void* f1, const void* g0, const void* g1,
                const void* o0, const void* o1, BcEntry* table_rw, double* eq_omega_state, cudaStream_t stream) {
  if constexpr (kExtCollision<COLL>) {
#if !XLBN_ON_HOST
    // D3Q19 with fp32 storage: the scalar tile kernel (step_tile.cuh) wherever the slab can be tiled, as for plain BGK
    if constexpr (L::Q == 19 && L::D == 3 && sizeof(TS) == 4) {
      const bool fits = tile1_eligible<L, TS>(p, f0, f1, g0, g1, o0 != nullptr || o1 != nullptr);
      if (requested_v == 501 && !fits)
        return fail(XLBN_E_SHAPE, "cells_per_thread = 501: the scalar tile kernel needs nz | 512, nz %% 16 == 0, ny %% (512 / nz) == 0, 16-byte aligned arrays and no halo handle (nz = %d, ny = %d)", p.nz, p.ny);
      // By default for SmagorinskyLESBGK in fp32 (512^3 cavity: 1.02 of the measured copy bandwidth against 0.93 for the direct-load
      // kernel); the forced operators only on request: the second equilibrium makes them issue-bound with the tile kernel's 16 warps
      // (forced BGK 0.835 against 0.876, forced Smagorinsky 0.825 against 0.822; profiles/r2_call19_matrix.txt), and the fp64 forced
      // Smagorinsky instantiation spills 330 bytes at 96 registers.
      if ((requested_v == 501 || (requested_v == 0 && sizeof(TC) == 4 && !kForcedCollision<COLL>)) && fits) return launch_step_tile1<L, COLL, TC, TS, 1>(p, x_count, stream);
    }
#endif
    return launch_step_v<L, COLL, TC, TS, 1>(p, x_count, stream);
  } else {
    return launch_step_base<L, COLL, TC, TS>(p, x_count, requested_v, f0, f1, g0, g1, o0, o1, table_rw, eq_omega_state, stream);
  }
}

// One entry per (lattice, collision); dispatches on (compute, store) dtype.  Defined in step_inst_*.cu.
struct StepCall {
  int compute_dtype, store_dtype, requested_v;
  const void* f0;
  void* f1;
  const uint8_t* bc;
  const uint32_t* miss;
  const BcEntry* table;
  BcEntry* table_rw;        // same table, writable (EquilibriumBC constants)
  double* eq_omega_state;  // host: omega the EquilibriumBC constants were computed for; NULL = stepper has no EquilibriumBC
  const uint8_t* kinds;  // host, 256 entries
  int nx, ny, nz, x_begin, x_count;
  double omega;
  const void* ghost_lo;
  const void* ghost_hi;
  void* out_lo;
  void* out_hi;
  cudaStream_t stream;
  double force[3];  // extended collision models only
  double smagorinsky;
  const double* eq_in;    // host: [4][kMaxQ] feq of the EquilibriumBC slots in the compute dtype (NULL: none)
  const uint8_t* eq_ids;  // host: [4]
};

template <class L, int COLL>
int dispatch_step(const StepCall& c);

// Host: feq(rho, u) of an EquilibriumBC in the compute dtype T, the expression of equilibrium<>() operation by operation (plain IEEE
// arithmetic: the host objects are compiled with -ffp-contract=off), stored as doubles (exact).  StepParams::eq_in.
template <class L, class T>
inline void equilibrium_on_host(double rho_d, const double* u_d, double* out) {
  const T rho = (T)rho_d;
  T u[L::D];
  for (int d = 0; d < L::D; ++d) u[d] = (T)u_d[d];
  T uu = u[0] * u[0];
  for (int d = 1; d < L::D; ++d) uu = uu + u[d] * u[d];
  const T usqr = T(1.5) * uu;
  static_for_host<L::Q>([&](auto l_) {
    constexpr int l = decltype(l_)::value;
    T cu = T(0);
    for (int d = 0; d < L::D; ++d) {
      if (L::c(d, l) == 1) cu += u[d];
      else if (L::c(d, l) == -1) cu -= u[d];
    }
    cu *= T(3.0);
    const T feq = rho * T(L::w(l)) * (T(1.0) + cu * (T(1.0) + T(0.5) * cu) - usqr);
    out[l] = (double)feq;
  });
}

// Host: fold everything that depends only on (population, x-plane class) into the kernel's pointer tables.
template <class L, class TS>
int fill_step_params(const StepCall& c, StepParams<TS>& p) {
  memset(&p, 0, sizeof(p));
  const TS* f0 = static_cast<const TS*>(c.f0);
  TS* f1 = static_cast<TS*>(c.f1);
  const TS* ghost_lo = static_cast<const TS*>(c.ghost_lo);
  const TS* ghost_hi = static_cast<const TS*>(c.ghost_hi);
  TS* out_lo = static_cast<TS*>(c.out_lo);
  TS* out_hi = static_cast<TS*>(c.out_hi);
  const long long plane = (long long)c.ny * c.nz;
  const long long n = plane * c.nx;
  if (n >= (1LL << 32)) return fail(XLBN_E_SHAPE, "more than 2^32 cells per slab (%lld): split the domain across GPUs", n);
  static_for_host<L::Q>([&](auto l_) {
    constexpr int l = decltype(l_)::value;
    constexpr int cx = L::ck(0, l);
    const TS* pop = f0 + (long long)l * n;
    // element offset inside the kernel is x*plane + row + z; the tables absorb the -cx*plane shift, ghosts and wraps
    p.pull[0][l] = pop - (long long)cx * plane;
    p.pull[1][l] = p.pull[0][l];
    p.pull[2][l] = p.pull[0][l];
    if (cx == 1)  // plane x = 0 pulls plane "x = -1"
      p.pull[1][l] = ghost_lo ? ghost_lo + (long long)L::xdir_slot(l) * plane : pop + (long long)(c.nx - 1) * plane;
    if (cx == -1)  // plane x = nx-1 pulls plane "x = nx"; the kernel adds (nx-1)*plane
      p.pull[2][l] = (ghost_hi ? ghost_hi + (long long)L::xdir_slot(l) * plane : pop) - (long long)(c.nx - 1) * plane;
    p.push[l] = f1 + (long long)l * n;
    if (cx == 1 && out_hi) p.peer_hi[l] = out_hi + (long long)L::xdir_slot(l) * plane - (long long)(c.nx - 1) * plane;
    if (cx == -1 && out_lo) p.peer_lo[l] = out_lo + (long long)L::xdir_slot(l) * plane;
  });
  p.bc = c.bc;
  memcpy(p.kinds, c.kinds, 256);
  p.f0 = f0;
  p.f1 = f1;
  p.f0w = const_cast<TS*>(f0);
  p.miss = c.miss;
  p.table = c.table;
  p.ghost_lo = ghost_lo;
  p.ghost_hi = ghost_hi;
  p.nx = c.nx;
  p.ny = c.ny;
  p.nz = c.nz;
  p.x_begin = c.x_begin;
  p.plane = plane;
  p.n = n;
  p.omega = c.omega;
  for (int a = 0; a < 3; ++a) p.force[a] = c.force[a];
  p.smagorinsky = c.smagorinsky;
  if (c.eq_in) {
    memcpy(p.eq_in, c.eq_in, sizeof(p.eq_in));
    memcpy(p.eq_ids, c.eq_ids, sizeof(p.eq_ids));
  }
  return 0;
}

template <class L, int COLL, class TC, class TS>
int run_step_typed(const StepCall& c) {
  StepParams<TS> p;
  if (int e = fill_step_params<L, TS>(c, p)) return e;
  return launch_step<L, COLL, TC, TS>(p, c.x_count, c.requested_v, c.f0, c.f1, c.ghost_lo, c.ghost_hi, c.out_lo, c.out_hi, c.table_rw, c.eq_omega_state, c.stream);
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
