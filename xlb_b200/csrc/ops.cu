// Stand-alone operator kernels (one thread per cell).  They exist for API / test parity with the reference's
// per-operator kernels and reuse the device functions of the fused step, so a stand-alone
// Stream -> Macroscopic -> Equilibrium -> Collision chain computes what the fused kernel computes.
// Arrays carry a runtime dtype; arithmetic is in the compute dtype TC.  Not the hot path.
#include "lbm_math.cuh"

namespace xlbn {

struct Dims {
  int nx, ny, nz;      // kernel extents (2-D fields are passed as (1, nx, ny))
  long long n;         // cells
};

static inline Dims kernel_dims(int lattice, const int32_t dims[3]) {
  Dims d;
  if (lattice == XLBN_D2Q9) {
    d.nx = 1;
    d.ny = dims[0];
    d.nz = dims[1];
  } else {
    d.nx = dims[0];
    d.ny = dims[1];
    d.nz = dims[2];
  }
  d.n = (long long)d.nx * d.ny * d.nz;
  return d;
}

static inline int check_dims(int lattice, const int32_t dims[3]) {
  if (!dims) return fail(XLBN_E_ARG, "dims is NULL");
  if (lattice < XLBN_D2Q9 || lattice > XLBN_D3Q27) return fail(XLBN_E_ARG, "unknown lattice %d", lattice);
  if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(XLBN_E_SHAPE, "non-positive dims %d %d %d", dims[0], dims[1], dims[2]);
  if (lattice == XLBN_D2Q9 && dims[2] != 1) return fail(XLBN_E_SHAPE, "2-D lattice needs dims[2] == 1, got %d", dims[2]);
  return 0;
}

#define XLBN_CELL_INDEX()                                             \
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; \
  if (i >= d.n) return;

static inline dim3 grid_for(long long n, int block = 256) { return dim3((unsigned)((n + block - 1) / block)); }

// ---- Stream (stream.py:59-99) ---------------------------------------------------------------------------------------
template <class L>
__global__ void stream_kernel(const void* fin, void* fout, int dtype, Dims d) {
  XLBN_CELL_INDEX();
  const int z = (int)(i % d.nz);
  const int y = (int)((i / d.nz) % d.ny);
  const int x = (int)(i / ((long long)d.nz * d.ny));
  XLBN_FOR(L::Q, l)
    int xs = x - L::ck(0, l), ys = y - L::ck(1, l), zs = z - L::ck(2, l);
    xs = xs < 0 ? d.nx - 1 : (xs >= d.nx ? 0 : xs);
    ys = ys < 0 ? d.ny - 1 : (ys >= d.ny ? 0 : ys);
    zs = zs < 0 ? d.nz - 1 : (zs >= d.nz ? 0 : zs);
    const long long src = (long long)l * d.n + ((long long)xs * d.ny + ys) * d.nz + zs;
    const long long dst = (long long)l * d.n + i;
    switch (dtype) {  // pure copy: no conversion round trip
      case XLBN_F16: reinterpret_cast<uint16_t*>(fout)[dst] = reinterpret_cast<const uint16_t*>(fin)[src]; break;
      case XLBN_F32: reinterpret_cast<uint32_t*>(fout)[dst] = reinterpret_cast<const uint32_t*>(fin)[src]; break;
      case XLBN_F64: reinterpret_cast<uint64_t*>(fout)[dst] = reinterpret_cast<const uint64_t*>(fin)[src]; break;
      default: reinterpret_cast<uint8_t*>(fout)[dst] = reinterpret_cast<const uint8_t*>(fin)[src]; break;  // bool / u8 masks
    }
  XLBN_END
}

// ---- QuadraticEquilibrium (quadratic_equilibrium.py:63-82) ------------------------------------------------------------
template <class L, class TC>
__global__ void equilibrium_kernel(const void* rho, int rho_dt, const void* u, int u_dt, void* f, int f_dt, Dims d) {
  XLBN_CELL_INDEX();
  TC uu[L::D], feq[L::Q];
  XLBN_FOR(L::D, a) uu[a] = load_as<TC>(u, u_dt, (long long)a * d.n + i); XLBN_END
  const TC r = load_as<TC>(rho, rho_dt, i);
  equilibrium<L, TC>(r, uu, feq);
  XLBN_FOR(L::Q, l) store_as<TC>(f, f_dt, (long long)l * d.n + i, feq[l]); XLBN_END
}

// ---- Macroscopic / moments (macroscopic.py:39-55, zero_moment.py:39-45, first_moment.py:40-60) -------------------------
template <class L, class TC>
__global__ void macroscopic_kernel(const void* f, int f_dt, void* rho, int rho_dt, void* u, int u_dt, Dims d) {
  XLBN_CELL_INDEX();
  TC ff[L::Q], r, uu[L::D];
  XLBN_FOR(L::Q, l) ff[l] = load_as<TC>(f, f_dt, (long long)l * d.n + i); XLBN_END
  macroscopic<L, TC>(ff, r, uu);
  if (rho) store_as<TC>(rho, rho_dt, i, r);
  if (u) {
    XLBN_FOR(L::D, a) store_as<TC>(u, u_dt, (long long)a * d.n + i, uu[a]); XLBN_END
  }
}

// FirstMoment with a caller-provided density: u = (sum c f) / rho  (first_moment.py:26-38)
template <class L, class TC>
__global__ void first_moment_kernel(const void* f, int f_dt, const void* rho, int rho_dt, void* u, int u_dt, Dims d) {
  XLBN_CELL_INDEX();
  TC uu[L::D];
  XLBN_FOR(L::D, a) uu[a] = TC(0); XLBN_END
  XLBN_FOR(L::Q, l)
    const TC fl = load_as<TC>(f, f_dt, (long long)l * d.n + i);
    XLBN_FOR(L::D, a)
      if constexpr (L::c(a, l) == 1) uu[a] += fl;
      else if constexpr (L::c(a, l) == -1) uu[a] -= fl;
    XLBN_END
  XLBN_END
  const TC r = load_as<TC>(rho, rho_dt, i);
  XLBN_FOR(L::D, a) store_as<TC>(u, u_dt, (long long)a * d.n + i, uu[a] / r); XLBN_END
}

template <class L, class TC>
__global__ void second_moment_kernel(const void* f, int f_dt, void* pi, int pi_dt, Dims d) {
  XLBN_CELL_INDEX();
  TC ff[L::Q], p[L::NT];
  XLBN_FOR(L::Q, l) ff[l] = load_as<TC>(f, f_dt, (long long)l * d.n + i); XLBN_END
  second_moment<L, TC>(ff, p);
  XLBN_FOR(L::NT, t) store_as<TC>(pi, pi_dt, (long long)t * d.n + i, p[t]); XLBN_END
}

// ---- Collision (bgk.py:37-62, kbc.py:299-329) --------------------------------------------------------------------------
template <class L, int COLL, class TC>
__global__ void collide_kernel(const void* f, int f_dt, const void* feq, int feq_dt, void* fout, int fout_dt, const void* rho, int rho_dt,
                               TC omega, Dims d) {
  XLBN_CELL_INDEX();
  TC ff[L::Q], fe[L::Q], out[L::Q];
  XLBN_FOR(L::Q, l)
    ff[l] = load_as<TC>(f, f_dt, (long long)l * d.n + i);
    fe[l] = load_as<TC>(feq, feq_dt, (long long)l * d.n + i);
  XLBN_END
  if constexpr (COLL == XLBN_BGK) {
    collide_bgk<L, TC>(ff, fe, omega, out);
  } else {
    const TC r = load_as<TC>(rho, rho_dt, i);
    collide_kbc<L, TC>(ff, fe, r, omega, out);
  }
  XLBN_FOR(L::Q, l) store_as<TC>(fout, fout_dt, (long long)l * d.n + i, out[l]); XLBN_END
}

// Any operator incl. SmagorinskyLESBGK and the ForcedCollision wrapper (smagorinsky_les_bgk.py:92-138, forced_collision.py:41-103):
// like the reference functionals it works on the GIVEN feq / rho / u, nothing is recomputed from f.
struct Force3 {
  double v[3];
};
template <class L, int COLL, class TC>
__global__ void collide_ext_kernel(const void* f, int f_dt, const void* feq, int feq_dt, void* fout, int fout_dt, const void* rho, int rho_dt,
                                   const void* u, int u_dt, TC omega, Force3 force, TC smagorinsky, Dims d) {
  XLBN_CELL_INDEX();
  TC ff[L::Q], fe[L::Q], out[L::Q];
  XLBN_FOR(L::Q, l)
    ff[l] = load_as<TC>(f, f_dt, (long long)l * d.n + i);
    fe[l] = load_as<TC>(feq, feq_dt, (long long)l * d.n + i);
  XLBN_END
  TC r = TC(1);
  if constexpr (kBaseCollision<COLL> == XLBN_KBC || kForcedCollision<COLL>) r = load_as<TC>(rho, rho_dt, i);
  if constexpr (kBaseCollision<COLL> == XLBN_BGK) collide_bgk<L, TC>(ff, fe, omega, out);
  else if constexpr (kBaseCollision<COLL> == XLBN_KBC) collide_kbc<L, TC>(ff, fe, r, omega, out);
  else if constexpr (L::D == 3) collide_smagorinsky<L, TC>(ff, fe, omega, smagorinsky, out);
  if constexpr (kForcedCollision<COLL>) {
    TC uu[L::D], fv[L::D];
    XLBN_FOR(L::D, a)
      uu[a] = load_as<TC>(u, u_dt, (long long)a * d.n + i);
      fv[a] = (TC)force.v[a];
    XLBN_END
    exact_difference<L, TC>(r, uu, fe, fv, out);
  }
  XLBN_FOR(L::Q, l) store_as<TC>(fout, fout_dt, (long long)l * d.n + i, out[l]); XLBN_END
}

// ExactDifference stand-alone (exact_difference_force.py:87-125)
template <class L, class TC>
__global__ void exact_difference_kernel(const void* fpc, int fpc_dt, const void* feq, int feq_dt, void* fout, int fout_dt, const void* rho, int rho_dt,
                                        const void* u, int u_dt, Force3 force, Dims d) {
  XLBN_CELL_INDEX();
  TC out[L::Q], fe[L::Q], uu[L::D], fv[L::D];
  XLBN_FOR(L::Q, l)
    out[l] = load_as<TC>(fpc, fpc_dt, (long long)l * d.n + i);
    fe[l] = load_as<TC>(feq, feq_dt, (long long)l * d.n + i);
  XLBN_END
  XLBN_FOR(L::D, a)
    uu[a] = load_as<TC>(u, u_dt, (long long)a * d.n + i);
    fv[a] = (TC)force.v[a];
  XLBN_END
  exact_difference<L, TC>(load_as<TC>(rho, rho_dt, i), uu, fe, fv, out);
  XLBN_FOR(L::Q, l) store_as<TC>(fout, fout_dt, (long long)l * d.n + i, out[l]); XLBN_END
}

// ---- Generic stand-alone boundary-condition kernel (boundary_condition.py:83-117) -------------------------------------
// f_0 := f_pre array, f_1 := f_post array, exactly as the reference passes them to the functional.
// bc_functional applies one BC's functional to one cell: f holds f_post on entry and the BC's result on exit.
template <class L, class TC>
__device__ __forceinline__ void bc_functional(const xlbn_bc_desc& bc, uint32_t miss, const void* fpre, int dt, TC aux, long long i, long long n,
                                              TC (&f)[L::Q]) {
  constexpr int Q = L::Q;
  const int kind = bc.kind;
  if (kind == XLBN_BC_EQUILIBRIUM) {
    TC u[L::D];
    XLBN_FOR(L::D, a) u[a] = (TC)bc.u[a]; XLBN_END
    equilibrium<L, TC>((TC)bc.rho, u, f);
  } else if (kind == XLBN_BC_FULLWAY_BOUNCE_BACK || bc_kind_needs_fpre(kind)) {
    TC pre[Q];
    XLBN_FOR(Q, l) pre[l] = load_as<TC>(fpre, dt, (long long)l * n + i); XLBN_END
    if (kind == XLBN_BC_FULLWAY_BOUNCE_BACK) {
      XLBN_FOR(Q, l) f[l] = pre[L::opp(l)]; XLBN_END
    } else if (kind == XLBN_BC_DO_NOTHING) {
      XLBN_FOR(Q, l) f[l] = pre[l]; XLBN_END
    } else {
      bc_take_opposite_of_pre<L, TC>(pre, miss, f);
    }
  } else if (bc_kind_needs_aux(kind)) {
    bc_zouhe<L, TC>(kind, aux, miss, f);
  }
}

template <class L, class TC>
__global__ void bc_apply_kernel(xlbn_bc_desc bc, const void* fpre, void* fpost, int dt, const uint8_t* bc_mask, const uint8_t* missing,
                                Dims d) {
  XLBN_CELL_INDEX();
  if (bc_mask[i] != (uint8_t)bc.id) return;  // other cells keep f_post
  constexpr int Q = L::Q;
  TC f[Q];
  uint32_t miss = 0;
  XLBN_FOR(Q, l)
    f[l] = load_as<TC>(fpost, dt, (long long)l * d.n + i);
    if (missing[(long long)l * d.n + i]) miss |= (1u << l);
  XLBN_END
  const TC aux = load_as<TC>(fpost, dt, i);  // f_1[0, cell]  (bc_zouhe.py:302, 327)
  bc_functional<L, TC>(bc, miss, fpre, dt, aux, i, d.n, f);
  XLBN_FOR(Q, l) store_as<TC>(fpost, dt, (long long)l * d.n + i, f[l]); XLBN_END
}

// ---- MomentumTransfer (force/momentum_transfer.py:108-176): momentum exchange over the edge cells of a no-slip BC -------
// edge cell: bc id matches and the rest direction is not missing.  f_post_stream = pull(f_0) with the BC's functional
// applied (f_pre = the cell's own f_0); m_d = sum over missing l of c[d, opp l] (f_0[opp l] + f_post_stream[l]).
template <class L, class TC>
__global__ void momentum_transfer_kernel(xlbn_bc_desc bc, const void* f0, const void* f1, int dt, const uint8_t* bc_mask,
                                         const uint8_t* missing, Dims d, double* force) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int Q = L::Q;
  double m[3] = {0.0, 0.0, 0.0};
  if (i < d.n && bc_mask[i] == (uint8_t)bc.id && !missing[i]) {
    const int z = (int)(i % d.nz);
    const int y = (int)((i / d.nz) % d.ny);
    const int x = (int)(i / ((long long)d.nz * d.ny));
    TC fpc[Q], fps[Q];
    uint32_t miss = 0;
    XLBN_FOR(Q, l)
      fpc[l] = load_as<TC>(f0, dt, (long long)l * d.n + i);
      if (missing[(long long)l * d.n + i]) miss |= (1u << l);
      int xs = x - L::ck(0, l), ys = y - L::ck(1, l), zs = z - L::ck(2, l);
      xs = xs < 0 ? d.nx - 1 : (xs >= d.nx ? 0 : xs);
      ys = ys < 0 ? d.ny - 1 : (ys >= d.ny ? 0 : ys);
      zs = zs < 0 ? d.nz - 1 : (zs >= d.nz ? 0 : zs);
      fps[l] = load_as<TC>(f0, dt, (long long)l * d.n + ((long long)xs * d.ny + ys) * d.nz + zs);
    XLBN_END
    const TC aux = f1 ? load_as<TC>(f1, dt, i) : TC(0);
    bc_functional<L, TC>(bc, miss, f0, dt, aux, i, d.n, fps);
    XLBN_FOR(Q, l)
      if ((miss >> l) & 1u) {
        const TC phi = fpc[L::opp(l)] + fps[l];
        XLBN_FOR(L::D, a)
          if constexpr (L::c(a, L::opp(l)) == 1) m[a] += (double)phi;
          else if constexpr (L::c(a, L::opp(l)) == -1) m[a] -= (double)phi;
        XLBN_END
      }
    XLBN_END
  }
  // warp reduction, then one atomic per warp and component
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double v = m[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(force + a, v);
  }
}

// ---- dispatch helpers ---------------------------------------------------------------------------------------------------
#define XLBN_LATTICE_SWITCH(lattice, ...)                            \
  switch (lattice) {                                                \
    case XLBN_D2Q9: { using L = D2Q9; __VA_ARGS__; } break;         \
    case XLBN_D3Q19: { using L = D3Q19; __VA_ARGS__; } break;       \
    case XLBN_D3Q27: { using L = D3Q27; __VA_ARGS__; } break;       \
    default: return fail(XLBN_E_ARG, "unknown lattice %d", lattice); \
  }

#define XLBN_COMPUTE_SWITCH(cdt, ...)                                                    \
  if ((cdt) == XLBN_F32) { using TC = float; __VA_ARGS__; }                              \
  else if ((cdt) == XLBN_F64) { using TC = double; __VA_ARGS__; }                        \
  else return fail(XLBN_E_DTYPE, "compute dtype must be F32 or F64, got %d", (int)(cdt));

#define XLBN_REQUIRE_FLOAT(dt, name) \
  if (!is_float_dtype(dt)) return fail(XLBN_E_DTYPE, "%s: dtype %d is not a floating type", name, (int)(dt));

}  // namespace xlbn

using namespace xlbn;

extern "C" {

int xlbn_stream(int lattice, const void* f_in, void* f_out, int dtype, const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_stream");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f_in || !f_out) return fail(XLBN_E_ARG, "xlbn_stream: NULL array");
  if (f_in == f_out) return fail(XLBN_E_ARG, "xlbn_stream: in-place streaming is not supported");
  if (dtype < XLBN_F16 || dtype > XLBN_BOOL) return fail(XLBN_E_DTYPE, "xlbn_stream: bad dtype %d", dtype);
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, stream_kernel<L><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(f_in, f_out, dtype, d));
  XLBN_LAUNCH_OK("stream_kernel");
  return 0;
}

int xlbn_equilibrium(int lattice, int compute_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype, void* f, int f_dtype,
                     const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_equilibrium");
  if (int e = check_dims(lattice, dims)) return e;
  if (!rho || !u || !f) return fail(XLBN_E_ARG, "xlbn_equilibrium: NULL array");
  XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
  XLBN_REQUIRE_FLOAT(u_dtype, "u");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, equilibrium_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      rho, rho_dtype, u, u_dtype, f, f_dtype, d)));
  XLBN_LAUNCH_OK("equilibrium_kernel");
  return 0;
}

int xlbn_macroscopic(int lattice, int compute_dtype, const void* f, int f_dtype, void* rho, int rho_dtype, void* u, int u_dtype,
                     const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_macroscopic");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f || (!rho && !u)) return fail(XLBN_E_ARG, "xlbn_macroscopic: NULL array");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  if (rho) XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
  if (u) XLBN_REQUIRE_FLOAT(u_dtype, "u");
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, macroscopic_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      f, f_dtype, rho, rho_dtype, u, u_dtype, d)));
  XLBN_LAUNCH_OK("macroscopic_kernel");
  return 0;
}

int xlbn_first_moment(int lattice, int compute_dtype, const void* f, int f_dtype, const void* rho, int rho_dtype, void* u, int u_dtype,
                      const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_first_moment");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f || !rho || !u) return fail(XLBN_E_ARG, "xlbn_first_moment: NULL array");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
  XLBN_REQUIRE_FLOAT(u_dtype, "u");
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, first_moment_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      f, f_dtype, rho, rho_dtype, u, u_dtype, d)));
  XLBN_LAUNCH_OK("first_moment_kernel");
  return 0;
}

int xlbn_second_moment(int lattice, int compute_dtype, const void* f, int f_dtype, void* pi, int pi_dtype, const int32_t dims[3],
                       void* stream) {
  XLBN_RANGE("xlbn_second_moment");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f || !pi) return fail(XLBN_E_ARG, "xlbn_second_moment: NULL array");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  XLBN_REQUIRE_FLOAT(pi_dtype, "pi");
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, second_moment_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      f, f_dtype, pi, pi_dtype, d)));
  XLBN_LAUNCH_OK("second_moment_kernel");
  return 0;
}

int xlbn_collide(int lattice, int collision, int compute_dtype, const void* f, int f_dtype, const void* feq, int feq_dtype, void* fout,
                 int fout_dtype, const void* rho, int rho_dtype, double omega, const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_collide");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f || !feq || !fout) return fail(XLBN_E_ARG, "xlbn_collide: NULL array");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  XLBN_REQUIRE_FLOAT(feq_dtype, "feq");
  XLBN_REQUIRE_FLOAT(fout_dtype, "fout");
  const Dims d = kernel_dims(lattice, dims);
  cudaStream_t st = (cudaStream_t)stream;
  if (collision == XLBN_BGK) {
    XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, collide_kernel<L, XLBN_BGK, TC><<<grid_for(d.n), 256, 0, st>>>(
                                                                        f, f_dtype, feq, feq_dtype, fout, fout_dtype, rho, rho_dtype, (TC)omega, d)));
  } else if (collision == XLBN_KBC) {
    if (lattice == XLBN_D3Q19) return fail(XLBN_E_UNSUPPORTED, "KBC: velocity set not supported: D3Q19 (reference: kbc.py:71-72, 184-185)");
    if (!rho) return fail(XLBN_E_ARG, "xlbn_collide: KBC needs rho");
    XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
    if (lattice == XLBN_D3Q27) {
      using L = D3Q27;
      XLBN_COMPUTE_SWITCH(compute_dtype, collide_kernel<L, XLBN_KBC, TC><<<grid_for(d.n), 256, 0, st>>>(f, f_dtype, feq, feq_dtype, fout, fout_dtype,
                                                                                                           rho, rho_dtype, (TC)omega, d));
    } else {
      using L = D2Q9;
      XLBN_COMPUTE_SWITCH(compute_dtype, collide_kernel<L, XLBN_KBC, TC><<<grid_for(d.n), 256, 0, st>>>(f, f_dtype, feq, feq_dtype, fout, fout_dtype,
                                                                                                           rho, rho_dtype, (TC)omega, d));
    }
  } else {
    return fail(XLBN_E_ARG, "unknown collision %d", collision);
  }
  XLBN_LAUNCH_OK("collide_kernel");
  return 0;
}

int xlbn_collide_ext(int lattice, int collision, int compute_dtype, const void* f, int f_dtype, const void* feq, int feq_dtype, void* fout,
                     int fout_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype, double omega, const double* force,
                     double smagorinsky, const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_collide_ext");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f || !feq || !fout) return fail(XLBN_E_ARG, "xlbn_collide_ext: NULL array");
  XLBN_REQUIRE_FLOAT(f_dtype, "f");
  XLBN_REQUIRE_FLOAT(feq_dtype, "feq");
  XLBN_REQUIRE_FLOAT(fout_dtype, "fout");
  const int base = collision & 3;
  const bool forced = (collision & XLBN_COLLISION_FORCED) != 0;
  if (collision < 0 || (collision & ~(3 | XLBN_COLLISION_FORCED)) || base == 3) return fail(XLBN_E_ARG, "unknown collision %d", collision);
  if (base == XLBN_KBC && lattice == XLBN_D3Q19) return fail(XLBN_E_UNSUPPORTED, "KBC: velocity set not supported: D3Q19 (reference: kbc.py:71-72, 184-185)");
  if (base == XLBN_SMAGORINSKY_LES_BGK && lattice == XLBN_D2Q9)
    return fail(XLBN_E_UNSUPPORTED, "SmagorinskyLESBGK: 3-D velocity sets only (reference: smagorinsky_les_bgk.py:71-76)");
  if (base == XLBN_KBC || forced) {
    if (!rho) return fail(XLBN_E_ARG, "xlbn_collide_ext: this operator needs rho");
    XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
  }
  Force3 fv = {{0.0, 0.0, 0.0}};
  if (forced) {
    if (!u || !force) return fail(XLBN_E_ARG, "xlbn_collide_ext: a forced operator needs u and the force vector");
    XLBN_REQUIRE_FLOAT(u_dtype, "u");
    for (int a = 0; a < (lattice == XLBN_D2Q9 ? 2 : 3); ++a) fv.v[a] = force[a];
  }
  const Dims d = kernel_dims(lattice, dims);
  cudaStream_t st = (cudaStream_t)stream;
#define XLBN_LAUNCH_COLLIDE_EXT(LAT, COLL)                                                                                                    \
  {                                                                                                                                           \
    using L = LAT;                                                                                                                            \
    XLBN_COMPUTE_SWITCH(compute_dtype, collide_ext_kernel<L, COLL, TC><<<grid_for(d.n), 256, 0, st>>>(                                         \
                                           f, f_dtype, feq, feq_dtype, fout, fout_dtype, rho, rho_dtype, u, u_dtype, (TC)omega, fv, (TC)smagorinsky, d)); \
  }
#define XLBN_COLLIDE_EXT_CASE(LAT, COLL)                                      \
  if (!forced) XLBN_LAUNCH_COLLIDE_EXT(LAT, COLL)                             \
  else XLBN_LAUNCH_COLLIDE_EXT(LAT, COLL | XLBN_COLLISION_FORCED)
  if (lattice == XLBN_D3Q19) {
    if (base == XLBN_BGK) { XLBN_COLLIDE_EXT_CASE(D3Q19, XLBN_BGK) }
    else { XLBN_COLLIDE_EXT_CASE(D3Q19, XLBN_SMAGORINSKY_LES_BGK) }
  } else if (lattice == XLBN_D3Q27) {
    if (base == XLBN_BGK) { XLBN_COLLIDE_EXT_CASE(D3Q27, XLBN_BGK) }
    else if (base == XLBN_KBC) { XLBN_COLLIDE_EXT_CASE(D3Q27, XLBN_KBC) }
    else { XLBN_COLLIDE_EXT_CASE(D3Q27, XLBN_SMAGORINSKY_LES_BGK) }
  } else {
    if (base == XLBN_BGK) { XLBN_COLLIDE_EXT_CASE(D2Q9, XLBN_BGK) }
    else { XLBN_COLLIDE_EXT_CASE(D2Q9, XLBN_KBC) }
  }
#undef XLBN_COLLIDE_EXT_CASE
#undef XLBN_LAUNCH_COLLIDE_EXT
  XLBN_LAUNCH_OK("collide_ext_kernel");
  return 0;
}

int xlbn_exact_difference(int lattice, int compute_dtype, const void* f_postcollision, int f_dtype, const void* feq, int feq_dtype, void* fout,
                          int fout_dtype, const void* rho, int rho_dtype, const void* u, int u_dtype, const double* force, const int32_t dims[3],
                          void* stream) {
  XLBN_RANGE("xlbn_exact_difference");
  if (int e = check_dims(lattice, dims)) return e;
  if (!f_postcollision || !feq || !fout || !rho || !u || !force) return fail(XLBN_E_ARG, "xlbn_exact_difference: NULL argument");
  XLBN_REQUIRE_FLOAT(f_dtype, "f_postcollision");
  XLBN_REQUIRE_FLOAT(feq_dtype, "feq");
  XLBN_REQUIRE_FLOAT(fout_dtype, "fout");
  XLBN_REQUIRE_FLOAT(rho_dtype, "rho");
  XLBN_REQUIRE_FLOAT(u_dtype, "u");
  Force3 fv = {{0.0, 0.0, 0.0}};
  for (int a = 0; a < (lattice == XLBN_D2Q9 ? 2 : 3); ++a) fv.v[a] = force[a];
  const Dims d = kernel_dims(lattice, dims);
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, exact_difference_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      f_postcollision, f_dtype, feq, feq_dtype, fout, fout_dtype, rho, rho_dtype, u, u_dtype, fv, d)));
  XLBN_LAUNCH_OK("exact_difference_kernel");
  return 0;
}

int xlbn_bc_apply(int lattice, int compute_dtype, const xlbn_bc_desc* bc, const void* f_pre, void* f_post, int dtype, const uint8_t* bc_mask,
                  const uint8_t* missing, const int32_t dims[3], void* stream) {
  XLBN_RANGE("xlbn_bc_apply");
  if (int e = check_dims(lattice, dims)) return e;
  if (!bc || !f_pre || !f_post || !bc_mask || !missing) return fail(XLBN_E_ARG, "xlbn_bc_apply: NULL argument");
  XLBN_REQUIRE_FLOAT(dtype, "f_pre/f_post");
  if (bc->kind <= XLBN_BC_NONE || bc->kind > XLBN_BC_EXTRAPOLATION_OUTFLOW) return fail(XLBN_E_ARG, "xlbn_bc_apply: bad BC kind %d", bc->kind);
  if (bc->id < 0 || bc->id > 255) return fail(XLBN_E_ARG, "xlbn_bc_apply: BC id %d outside uint8", bc->id);
  const Dims d = kernel_dims(lattice, dims);
  const xlbn_bc_desc b = *bc;
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, bc_apply_kernel<L, TC><<<grid_for(d.n), 256, 0, (cudaStream_t)stream>>>(
                                                                      b, f_pre, f_post, dtype, bc_mask, missing, d)));
  XLBN_LAUNCH_OK("bc_apply_kernel");
  return 0;
}

int xlbn_momentum_transfer(int lattice, int compute_dtype, const xlbn_bc_desc* bc, const void* f0, const void* f1, int dtype,
                           const uint8_t* bc_mask, const uint8_t* missing, const int32_t dims[3], double* force, void* stream) {
  XLBN_RANGE("xlbn_momentum_transfer");
  if (int e = check_dims(lattice, dims)) return e;
  if (!bc || !f0 || !bc_mask || !missing || !force) return fail(XLBN_E_ARG, "xlbn_momentum_transfer: NULL argument");
  XLBN_REQUIRE_FLOAT(dtype, "f_0");
  if (bc->kind <= XLBN_BC_NONE || bc->kind > XLBN_BC_EXTRAPOLATION_OUTFLOW) return fail(XLBN_E_ARG, "xlbn_momentum_transfer: bad BC kind %d", bc->kind);
  const Dims d = kernel_dims(lattice, dims);
  const xlbn_bc_desc b = *bc;
  cudaStream_t st = (cudaStream_t)stream;
  XLBN_CUDA_OK(cudaMemsetAsync(force, 0, 3 * sizeof(double), st));
  XLBN_LATTICE_SWITCH(lattice, XLBN_COMPUTE_SWITCH(compute_dtype, momentum_transfer_kernel<L, TC><<<grid_for(d.n), 256, 0, st>>>(
                                                                      b, f0, f1, dtype, bc_mask, missing, d, force)));
  XLBN_LAUNCH_OK("momentum_transfer_kernel");
  return 0;
}

}  // extern "C"
