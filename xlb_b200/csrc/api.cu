// extern "C" entry points of the hot path: stepper handle + one fused step (include/xlb_b200.h).
#include <cmath>

#include "halo.cuh"
#include "step_kernel.cuh"

struct xlbn_stepper {
  int lattice, collision, compute_dtype, store_dtype, cells_per_thread;
  bool needs_missing;
  uint8_t kinds[256];    // host copy: bc id -> kind
  bool has_equilibrium_bc;
  double eq_omega;       // omega the EquilibriumBC constants in the table were computed for (NaN = never)
  xlbn::BcEntry* table;  // device, 256 entries
  int device;
  double eq_in[4][xlbn::kMaxQ];  // feq(rho, u) of the first 4 EquilibriumBC ids in the compute dtype (StepParams::eq_in)
  uint8_t eq_ids[4];
  bool forced;           // ForcedCollision (xlbn_stepper_set_force)
  double force[3];
  double smagorinsky;
};

using namespace xlbn;

namespace {
// Makes `device` current for the duration of a call when it is not already (one cudaGetDevice per call otherwise).
struct DeviceGuard {
  int previous = -1;
  cudaError_t error = cudaSuccess;
  explicit DeviceGuard(int device) {
    int current = -1;
    error = cudaGetDevice(&current);
    if (error == cudaSuccess && current != device) {
      error = cudaSetDevice(device);
      if (error == cudaSuccess) previous = current;
    }
  }
  ~DeviceGuard() {
    if (previous >= 0) cudaSetDevice(previous);
  }
};
}  // namespace

namespace xlbn {
template <> int dispatch_step<D3Q19, XLBN_BGK>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_BGK>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_KBC>(const StepCall&);
template <> int dispatch_step<D2Q9, XLBN_BGK>(const StepCall&);
template <> int dispatch_step<D2Q9, XLBN_KBC>(const StepCall&);
// extended collision models (step_inst_ext_*.cu)
constexpr int kF = XLBN_COLLISION_FORCED;
template <> int dispatch_step<D3Q19, XLBN_BGK | kF>(const StepCall&);
template <> int dispatch_step<D3Q19, XLBN_SMAGORINSKY_LES_BGK>(const StepCall&);
template <> int dispatch_step<D3Q19, XLBN_SMAGORINSKY_LES_BGK | kF>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_BGK | kF>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_KBC | kF>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_SMAGORINSKY_LES_BGK>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_SMAGORINSKY_LES_BGK | kF>(const StepCall&);
template <> int dispatch_step<D2Q9, XLBN_BGK | kF>(const StepCall&);
template <> int dispatch_step<D2Q9, XLBN_KBC | kF>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_KBC | kLeanKbc>(const StepCall&);  // tuning variant (cells_per_thread = 301)
template <> int dispatch_step<D2Q9, XLBN_KBC | kLeanKbc>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_KBC | kExactKbc>(const StepCall&);  // parity form (cells_per_thread = 300)
template <> int dispatch_step<D2Q9, XLBN_KBC | kExactKbc>(const StepCall&);
template <> int dispatch_step<D2Q9X, XLBN_BGK>(const StepCall&);  // 2-D x-slab axis order (step_inst_d2q9x.cu)
template <> int dispatch_step<D2Q9X, XLBN_KBC>(const StepCall&);
template <> int dispatch_step<D2Q9X, XLBN_KBC | kLeanKbc>(const StepCall&);
template <> int dispatch_step<D3Q27, XLBN_KBC | kLeanKbc | kF>(const StepCall&);
template <> int dispatch_step<D2Q9, XLBN_KBC | kLeanKbc | kF>(const StepCall&);
}  // namespace xlbn

extern "C" {

int xlbn_version(void) { return XLBN_VERSION; }

const char* xlbn_last_error(void) { return error_buffer(); }

int xlbn_lattice_tables(int lattice, int32_t* c, double* w, int32_t* opp) {
  auto fill = [&](auto tag) {
    using L = decltype(tag);
    for (int l = 0; l < L::Q; ++l) {
      if (c)
        for (int a = 0; a < 3; ++a) c[a * L::Q + l] = a < L::D ? L::c(a, l) : 0;
      if (w) w[l] = L::w(l);
      if (opp) opp[l] = L::opp(l);
    }
    return (int)L::Q;
  };
  switch (lattice) {
    case XLBN_D2Q9: return fill(D2Q9{});
    case XLBN_D3Q19: return fill(D3Q19{});
    case XLBN_D3Q27: return fill(D3Q27{});
    default: return fail(XLBN_E_ARG, "unknown lattice %d", lattice);
  }
}

int xlbn_stepper_create(const xlbn_stepper_desc* desc, xlbn_stepper** out) {
  XLBN_RANGE("xlbn_stepper_create");
  if (!desc || !out) return fail(XLBN_E_ARG, "stepper_create: NULL argument");
  if (desc->lattice < XLBN_D2Q9 || desc->lattice > XLBN_D3Q27) return fail(XLBN_E_ARG, "stepper_create: unknown lattice %d", desc->lattice);
  if (desc->collision != XLBN_BGK && desc->collision != XLBN_KBC && desc->collision != XLBN_SMAGORINSKY_LES_BGK)
    return fail(XLBN_E_ARG, "stepper_create: unknown collision %d", desc->collision);
  if (desc->collision == XLBN_SMAGORINSKY_LES_BGK && desc->lattice == XLBN_D2Q9)
    return fail(XLBN_E_UNSUPPORTED, "SmagorinskyLESBGK: 3-D velocity sets only (the reference functional reads c[2, l]: smagorinsky_les_bgk.py:71-76)");
  if (desc->collision == XLBN_KBC && desc->lattice == XLBN_D3Q19)
    return fail(XLBN_E_UNSUPPORTED, "KBC: velocity set not supported: D3Q19 (reference: kbc.py:71-72, 184-185)");
  if (desc->compute_dtype != XLBN_F32 && desc->compute_dtype != XLBN_F64) return fail(XLBN_E_DTYPE, "stepper_create: compute dtype %d", desc->compute_dtype);
  if (!is_float_dtype(desc->store_dtype)) return fail(XLBN_E_DTYPE, "stepper_create: store dtype %d", desc->store_dtype);
  if (desc->compute_dtype == XLBN_F32 && desc->store_dtype == XLBN_F64) return fail(XLBN_E_DTYPE, "stepper_create: no FP32FP64 policy");
  if (desc->n_bc < 0 || (desc->n_bc > 0 && !desc->bcs)) return fail(XLBN_E_ARG, "stepper_create: bad BC list");
  const int cpt = desc->cells_per_thread;
  if (cpt != 0 && cpt != 1 && cpt != 2 && cpt != 4 && cpt != 8 && cpt != 102 && cpt != 104 && cpt != 202 && cpt != 203 && cpt != 300 && cpt != 301 && cpt != 402 && cpt != 403 && cpt != 404 && cpt != 501 && cpt != 502)
    return fail(XLBN_E_ARG, "stepper_create: cells_per_thread = %d", cpt);
  if ((cpt == 300 || cpt == 301) && desc->collision != XLBN_KBC)
    return fail(XLBN_E_ARG, "stepper_create: cells_per_thread = %d selects a KBC formulation; the stepper's collision is %d", cpt, desc->collision);

  BcEntry host[256];
  memset(host, 0, sizeof(host));
  bool needs_missing = false;
  for (int i = 0; i < desc->n_bc; ++i) {
    const xlbn_bc_desc& b = desc->bcs[i];
    if (b.id < 1 || b.id > 254) return fail(XLBN_E_ARG, "stepper_create: BC id %d outside 1..254", b.id);
    if (b.kind <= XLBN_BC_NONE || b.kind > XLBN_BC_EXTRAPOLATION_OUTFLOW) return fail(XLBN_E_ARG, "stepper_create: BC kind %d", b.kind);
    if (host[b.id].kind != 0) return fail(XLBN_E_ARG, "stepper_create: duplicate BC id %d", b.id);
    host[b.id].kind = b.kind;
    host[b.id].rho = b.rho;
    for (int a = 0; a < 3; ++a) host[b.id].u[a] = b.u[a];
    needs_missing |= bc_kind_needs_missing(b.kind);
  }
  xlbn_stepper* s = new xlbn_stepper();
  s->lattice = desc->lattice;
  s->collision = desc->collision;
  s->compute_dtype = desc->compute_dtype;
  s->store_dtype = desc->store_dtype;
  s->cells_per_thread = cpt;
  s->needs_missing = needs_missing;
  s->table = nullptr;
  s->has_equilibrium_bc = false;
  for (int i = 0; i < 256; ++i) {
    s->kinds[i] = (uint8_t)host[i].kind;
    s->has_equilibrium_bc |= host[i].kind == XLBN_BC_EQUILIBRIUM;
  }
  s->eq_omega = nan("");
  memset(s->eq_in, 0, sizeof(s->eq_in));
  memset(s->eq_ids, 0, sizeof(s->eq_ids));
  for (int id = 1, n = 0; id < 255 && n < 4; ++id) {
    if (host[id].kind != XLBN_BC_EQUILIBRIUM) continue;
    auto fill = [&](auto lat) {
      using L = decltype(lat);
      if (desc->compute_dtype == XLBN_F32) equilibrium_on_host<L, float>(host[id].rho, host[id].u, s->eq_in[n]);
      else equilibrium_on_host<L, double>(host[id].rho, host[id].u, s->eq_in[n]);
    };
    if (desc->lattice == XLBN_D2Q9) fill(D2Q9{});
    else if (desc->lattice == XLBN_D3Q19) fill(D3Q19{});
    else fill(D3Q27{});
    s->eq_ids[n++] = (uint8_t)id;
  }
  s->forced = false;
  s->force[0] = s->force[1] = s->force[2] = 0.0;
  s->smagorinsky = 0.17;  // smagorinsky_les_bgk.py:24
  cudaError_t e = cudaGetDevice(&s->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->table), sizeof(host));
  if (e == cudaSuccess) e = cudaMemcpy(s->table, host, sizeof(host), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (s->table) cudaFree(s->table);
    delete s;
    return cuda_fail(e, "stepper_create");
  }
  *out = s;
  return 0;
}

int xlbn_stepper_destroy(xlbn_stepper* s) {
  XLBN_RANGE("xlbn_stepper_destroy");
  if (!s) return 0;
  if (s->table) cudaFree(s->table);
  delete s;
  return 0;
}

int xlbn_stepper_set_force(xlbn_stepper* s, const double* force) {
  if (!s) return fail(XLBN_E_ARG, "stepper_set_force: NULL stepper");
  s->forced = force != nullptr;
  for (int a = 0; a < 3; ++a) s->force[a] = (force && a < (s->lattice == XLBN_D2Q9 ? 2 : 3)) ? force[a] : 0.0;
  return 0;
}

int xlbn_stepper_set_smagorinsky(xlbn_stepper* s, double coefficient) {
  if (!s) return fail(XLBN_E_ARG, "stepper_set_smagorinsky: NULL stepper");
  if (s->collision != XLBN_SMAGORINSKY_LES_BGK) return fail(XLBN_E_STATE, "stepper_set_smagorinsky: the stepper's collision is not SmagorinskyLESBGK");
  if (!(coefficient >= 0.0)) return fail(XLBN_E_ARG, "stepper_set_smagorinsky: coefficient %g", coefficient);
  s->smagorinsky = coefficient;
  return 0;
}

int xlbn_stepper_prepare(xlbn_stepper* s, double omega, void* stream) {
  XLBN_RANGE("xlbn_stepper_prepare");
  if (!s) return fail(XLBN_E_ARG, "stepper_prepare: NULL stepper");
  // only the FP32FP16 BGK pair path reads per-omega constants (BcEntry::eq_out of EquilibriumBC entries)
  if (!s->has_equilibrium_bc || s->compute_dtype != XLBN_F32 || s->store_dtype != XLBN_F16 || s->collision != XLBN_BGK || s->forced) return 0;
  DeviceGuard guard(s->device);
  if (guard.error != cudaSuccess) return cuda_fail(guard.error, "stepper_prepare: cudaSetDevice");
  cudaStream_t st = (cudaStream_t)stream;
  switch (s->lattice) {
    case XLBN_D2Q9: bc_precompute_kernel<D2Q9, XLBN_BGK><<<1, 256, 0, st>>>(s->table, (float)omega); break;
    case XLBN_D3Q19: bc_precompute_kernel<D3Q19, XLBN_BGK><<<1, 256, 0, st>>>(s->table, (float)omega); break;
    default: bc_precompute_kernel<D3Q27, XLBN_BGK><<<1, 256, 0, st>>>(s->table, (float)omega); break;
  }
  XLBN_LAUNCH_OK("bc_precompute_kernel");
  s->eq_omega = omega;
  return 0;
}

int xlbn_step(xlbn_stepper* s, const void* f0, void* f1, const uint8_t* bc_mask, const uint32_t* missing_bits, const xlbn_domain* dom, double omega,
              int timestep, xlbn_halo* halo, void* stream) {
  XLBN_RANGE("xlbn_step");
  if (!s || !f0 || !f1 || !bc_mask || !dom) return fail(XLBN_E_ARG, "xlbn_step: NULL argument");
  if (f0 == f1) return fail(XLBN_E_ARG, "xlbn_step: f0 and f1 must be different buffers (pull scheme)");
  if (s->needs_missing && !missing_bits) return fail(XLBN_E_ARG, "xlbn_step: this stepper has BCs that read the missing-direction bitmask");
  if (dom->nx <= 0 || dom->ny <= 0 || dom->nz <= 0) return fail(XLBN_E_SHAPE, "xlbn_step: dims %d %d %d", dom->nx, dom->ny, dom->nz);
  if (dom->x_begin < 0 || dom->x_count < 0 || dom->x_begin + dom->x_count > dom->nx)
    return fail(XLBN_E_SHAPE, "xlbn_step: x range [%d, %d) outside [0, %d)", dom->x_begin, dom->x_begin + dom->x_count, dom->nx);
  if (s->lattice == XLBN_D2Q9 && dom->nz != 1) return fail(XLBN_E_SHAPE, "xlbn_step: 2-D lattice needs nz == 1");
  if (dom->x_count == 0) return 0;
  DeviceGuard guard(s->device);  // the BC table lives on the device the stepper was created on; launch there
  if (guard.error != cudaSuccess) return cuda_fail(guard.error, "xlbn_step: cudaSetDevice");

  StepCall c;
  c.compute_dtype = s->compute_dtype;
  c.store_dtype = s->store_dtype;
  c.requested_v = s->cells_per_thread == 300 ? 1 : s->cells_per_thread;
  c.f0 = f0;
  c.f1 = f1;
  c.bc = bc_mask;
  c.miss = missing_bits;
  c.table = s->table;
  c.table_rw = s->table;
  // EquilibriumBC constants (read by the FP32FP16 pair path only) follow omega; refreshed stream-ordered when it changes
  c.eq_omega_state = s->has_equilibrium_bc ? &s->eq_omega : nullptr;
  c.kinds = s->kinds;
  c.omega = omega;
  c.stream = (cudaStream_t)stream;
  c.ghost_lo = c.ghost_hi = nullptr;
  c.out_lo = c.out_hi = nullptr;
  for (int a = 0; a < 3; ++a) c.force[a] = s->force[a];  // physical components; the cell algebra uses L::c(), not kernel axes
  c.smagorinsky = s->smagorinsky;
  c.eq_in = &s->eq_in[0][0];
  c.eq_ids = s->eq_ids;
  const bool slab_2d = s->lattice == XLBN_D2Q9 && (halo || dom->x_begin != 0 || dom->x_count != dom->nx);
  if (slab_2d) {  // x-slab / partial x range in 2-D: kernel extents (nx, 1, ny), physical x on the kernel's slab axis (D2Q9X)
    if (s->forced) return fail(XLBN_E_UNSUPPORTED, "xlbn_step: forced collision on a 2-D slab / partial x range is not built");
    c.nx = dom->nx;
    c.ny = 1;
    c.nz = dom->ny;
    c.x_begin = dom->x_begin;
    c.x_count = dom->x_count;
  } else if (s->lattice == XLBN_D2Q9) {  // run [q][nx][ny] as kernel extents (1, nx, ny): unit-stride axis = thread axis
    c.nx = 1;
    c.ny = dom->nx;
    c.nz = dom->ny;
    c.x_begin = 0;
    c.x_count = 1;
  } else {
    c.nx = dom->nx;
    c.ny = dom->ny;
    c.nz = dom->nz;
    c.x_begin = dom->x_begin;
    c.x_count = dom->x_count;
  }
  if (halo) {
    if (!halo->connected) return fail(XLBN_E_STATE, "xlbn_step: halo is not connected");
    if (int e = halo_check_alive(halo, "xlbn_step")) return e;
    if (halo->lattice != s->lattice || halo->store_dtype != s->store_dtype || halo->ny != dom->ny || halo->nz != dom->nz)
      return fail(XLBN_E_SHAPE, "xlbn_step: halo does not match the stepper / domain");
    const int p_in = timestep & 1, p_out = (timestep + 1) & 1;
    c.ghost_lo = halo_ghost(halo, halo->base, p_in, 0);
    c.ghost_hi = halo_ghost(halo, halo->base, p_in, 1);
    c.out_hi = halo_ghost(halo, halo->peer_hi, p_out, 0);
    c.out_lo = halo_ghost(halo, halo->peer_lo, p_out, 1);
  }
  // KBC: the register-lean formulation is the default (B200: 0.81 vs 0.69 of the HBM roofline, profiles/r2_*); 300 = literal
  // (forced KBC: the lean form too, except in the 2-D slab axis order, where the forced operators are not built)
  const bool lean = s->collision == XLBN_KBC && (s->cells_per_thread == 0 || s->cells_per_thread == 301) && !(s->forced && slab_2d);
  const bool exact_kbc = s->collision == XLBN_KBC && !s->forced && s->cells_per_thread == 300 && !slab_2d;
  const int coll = s->collision | (s->forced ? kF : 0) | (lean ? kLeanKbc : 0) | (exact_kbc ? kExactKbc : 0);
  switch (s->lattice) {
    case XLBN_D3Q19:
      switch (coll) {
        case XLBN_BGK: return dispatch_step<D3Q19, XLBN_BGK>(c);
        case XLBN_BGK | kF: return dispatch_step<D3Q19, XLBN_BGK | kF>(c);
        case XLBN_SMAGORINSKY_LES_BGK: return dispatch_step<D3Q19, XLBN_SMAGORINSKY_LES_BGK>(c);
        case XLBN_SMAGORINSKY_LES_BGK | kF: return dispatch_step<D3Q19, XLBN_SMAGORINSKY_LES_BGK | kF>(c);
      }
      break;
    case XLBN_D3Q27:
      switch (coll) {
        case XLBN_BGK: return dispatch_step<D3Q27, XLBN_BGK>(c);
        case XLBN_KBC: return dispatch_step<D3Q27, XLBN_KBC>(c);
        case XLBN_BGK | kF: return dispatch_step<D3Q27, XLBN_BGK | kF>(c);
        case XLBN_KBC | kF: return dispatch_step<D3Q27, XLBN_KBC | kF>(c);
        case XLBN_KBC | kLeanKbc: return dispatch_step<D3Q27, XLBN_KBC | kLeanKbc>(c);
        case XLBN_KBC | kLeanKbc | kF: return dispatch_step<D3Q27, XLBN_KBC | kLeanKbc | kF>(c);
        case XLBN_KBC | kExactKbc: return dispatch_step<D3Q27, XLBN_KBC | kExactKbc>(c);
        case XLBN_SMAGORINSKY_LES_BGK: return dispatch_step<D3Q27, XLBN_SMAGORINSKY_LES_BGK>(c);
        case XLBN_SMAGORINSKY_LES_BGK | kF: return dispatch_step<D3Q27, XLBN_SMAGORINSKY_LES_BGK | kF>(c);
      }
      break;
    case XLBN_D2Q9:
      if (slab_2d) {
        switch (coll) {
          case XLBN_BGK: return dispatch_step<D2Q9X, XLBN_BGK>(c);
          case XLBN_KBC: return dispatch_step<D2Q9X, XLBN_KBC>(c);
          case XLBN_KBC | kLeanKbc: return dispatch_step<D2Q9X, XLBN_KBC | kLeanKbc>(c);
        }
        break;
      }
      switch (coll) {
        case XLBN_BGK: return dispatch_step<D2Q9, XLBN_BGK>(c);
        case XLBN_KBC: return dispatch_step<D2Q9, XLBN_KBC>(c);
        case XLBN_BGK | kF: return dispatch_step<D2Q9, XLBN_BGK | kF>(c);
        case XLBN_KBC | kF: return dispatch_step<D2Q9, XLBN_KBC | kF>(c);
        case XLBN_KBC | kLeanKbc: return dispatch_step<D2Q9, XLBN_KBC | kLeanKbc>(c);
        case XLBN_KBC | kLeanKbc | kF: return dispatch_step<D2Q9, XLBN_KBC | kLeanKbc | kF>(c);
        case XLBN_KBC | kExactKbc: return dispatch_step<D2Q9, XLBN_KBC | kExactKbc>(c);
      }
      break;
  }
  return fail(XLBN_E_ARG, "xlbn_step: bad stepper (lattice %d, collision %d)", s->lattice, coll);
}

}  // extern "C"
