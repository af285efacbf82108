// x-slab halo: compact ghost planes + step counters of one slab, peer-mapped into its two ring neighbours.
// Replaces the two lax.ppermute collectives of the reference (xlb/distribute/distribute.py:23-44): instead of fixing up
// wrongly wrapped planes AFTER streaming, the outgoing populations of the slab faces are written into the neighbours'
// ghost planes BEFORE the next pull — by the step kernel itself through peer pointers (step_kernel.cuh), or by
// xlbn_halo_push for the initial fill / non-fused fallback.  Same data, same ring (periodic in x across ranks).
//
// Ghost block (one cudaMalloc, exportable with CUDA IPC):
//   [parity 0|1][face 0|1][n_dir][ny][nz] store dtype, then (256-B aligned) int flags[4].
// Protocol for step t (reads state t, produces state t+1):
//   wait(t)  : both neighbours have delivered their state-t face populations into my parity-(t&1) ghosts
//   step     : pulls read parity t&1; my new face populations go into the neighbours' parity-((t+1)&1) ghosts
//   signal(t+1)
// Double buffering is sufficient: a neighbour overwrites my parity-p ghosts for state t+2 only after it has seen my
// signal(t+1), which I send after my face planes of step t (the only readers of state-t ghosts) are done.
#include "halo.cuh"
#include "lbm_math.cuh"

#include <cstdlib>
#include <cstring>

namespace xlbn {

static int lattice_ndir(int lattice) {
  switch (lattice) {
    case XLBN_D3Q19: return D3Q19::n_xdir();
    case XLBN_D3Q27: return D3Q27::n_xdir();
    case XLBN_D2Q9: return D2Q9X::n_xdir();  // 2-D slabs run in the D2Q9X axis order (lattice.cuh)
    default: return -1;
  }
}

// plane -> peer ghost copy of the populations with ck(0) == sx (initial fill / non-fused fallback)
template <class L>
__global__ void halo_push_kernel(const char* f, char* dst, int sx, int x_plane, long long plane, long long n, int esize) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // element within the plane
  if (i >= plane) return;
  XLBN_FOR(L::Q, l)
    if constexpr (L::ck(0, l) != 0) {
      if (L::ck(0, l) == sx) {
        const long long src = ((long long)l * n + (long long)x_plane * plane + i) * esize;
        const long long d = ((long long)L::xdir_slot(l) * plane + i) * esize;
        for (int b = 0; b < esize; ++b) dst[d + b] = f[src + b];
      }
    }
  XLBN_END
}

__global__ void halo_signal_kernel(int* flag_a, int* flag_b, int value) {
  __threadfence_system();  // everything this GPU stored to the peers before this kernel is visible before the flag
  *reinterpret_cast<volatile int*>(flag_a) = value;
  *reinterpret_cast<volatile int*>(flag_b) = value;
  __threadfence_system();
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Never hangs the GPU: after timeout_ns the wait gives up, marks the handle dead (flags[2] on the device and *timed_out in
// mapped host memory, value = step + 1) and lets the stream drain; the host side then refuses every further call.
__global__ void halo_wait_kernel(int* flags, int value, long long timeout_ns, int* timed_out) {
  volatile int* f = flags;
  const long long t0 = global_ns();
  while (f[0] < value || f[1] < value) {
    if (global_ns() - t0 > timeout_ns) {
      f[2] = value + 1;
      *reinterpret_cast<volatile int*>(timed_out) = value + 1;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

int halo_check_alive(const xlbn_halo* h, const char* where) {
  if (h->timed_out && *h->timed_out != 0)
    return fail(XLBN_E_STATE,
                "%s: a ring neighbour did not deliver its face populations for step %d within %.0f s (dead or stalled rank); the ghost planes "
                "read since then were stale, so the populations of this slab are invalid from that step on",
                where, *h->timed_out - 1, (double)h->timeout_ns * 1e-9);
  return 0;
}

}  // namespace xlbn

using namespace xlbn;

extern "C" {

int xlbn_halo_create(int lattice, int store_dtype, int ny, int nz, xlbn_halo** out) {
  XLBN_RANGE("xlbn_halo_create");
  if (!out) return fail(XLBN_E_ARG, "halo_create: out is NULL");
  const int ndir = lattice_ndir(lattice);
  if (ndir < 0) return fail(XLBN_E_UNSUPPORTED, "halo: unknown lattice %d", lattice);
  if (!is_float_dtype(store_dtype)) return fail(XLBN_E_DTYPE, "halo: bad store dtype %d", store_dtype);
  if (ny <= 0 || nz <= 0) return fail(XLBN_E_SHAPE, "halo: ny=%d nz=%d", ny, nz);
  xlbn_halo* h = new xlbn_halo();
  h->lattice = lattice;
  h->store_dtype = store_dtype;
  h->ny = ny;
  h->nz = nz;
  h->ndir = ndir;
  h->plane_bytes = (size_t)ny * nz * dtype_size(store_dtype);
  h->flags_offset = ((4 * (size_t)ndir * h->plane_bytes + 255) / 256) * 256;
  h->block_bytes = h->flags_offset + 256;
  h->peer_lo = h->peer_hi = nullptr;
  h->ipc_lo = h->ipc_hi = h->connected = false;
  h->base = nullptr;
  h->timed_out = nullptr;
  h->timed_out_dev = nullptr;
  h->timeout_ns = 120LL * 1000000000LL;  // default 120 s; XLBN_HALO_TIMEOUT_S or xlbn_halo_set_timeout change it
  if (const char* env = getenv("XLBN_HALO_TIMEOUT_S")) {
    const double sec = atof(env);
    if (sec > 0.0) h->timeout_ns = (long long)(sec * 1e9);
  }
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(const_cast<int**>(&h->timed_out)), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *h->timed_out = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->timed_out_dev), const_cast<int*>(h->timed_out), 0);
  }
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->base), h->block_bytes);
  if (e == cudaSuccess) e = cudaMemset(h->base, 0, h->block_bytes);
  if (e == cudaSuccess) {
    const int init[4] = {-1, -1, 0, 0};
    e = cudaMemcpy(halo_flags(h, h->base), init, sizeof(init), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    if (h->base) cudaFree(h->base);
    if (h->timed_out) cudaFreeHost(const_cast<int*>(h->timed_out));
    delete h;
    return cuda_fail(e, "halo_create");
  }
  *out = h;
  return 0;
}

int xlbn_halo_destroy(xlbn_halo* h) {
  XLBN_RANGE("xlbn_halo_destroy");
  if (!h) return 0;
  if (h->ipc_lo && h->peer_lo) cudaIpcCloseMemHandle(h->peer_lo);
  if (h->ipc_hi && h->peer_hi && h->peer_hi != h->peer_lo) cudaIpcCloseMemHandle(h->peer_hi);
  if (h->base) cudaFree(h->base);
  if (h->timed_out) cudaFreeHost(const_cast<int*>(h->timed_out));
  delete h;
  return 0;
}

int xlbn_halo_set_timeout(xlbn_halo* h, double seconds) {
  if (!h) return fail(XLBN_E_ARG, "halo_set_timeout: NULL");
  if (!(seconds > 0.0)) return fail(XLBN_E_ARG, "halo_set_timeout: %g s", seconds);
  h->timeout_ns = (long long)(seconds * 1e9);
  return 0;
}

int xlbn_halo_timed_out(xlbn_halo* h) {
  if (!h) return fail(XLBN_E_ARG, "halo_timed_out: NULL");
  return (h->timed_out && *h->timed_out != 0) ? 1 : 0;
}

int xlbn_halo_export(xlbn_halo* h, unsigned char handle[64]) {
  XLBN_RANGE("xlbn_halo_export");
  if (!h || !handle) return fail(XLBN_E_ARG, "halo_export: NULL");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t ipc;
  XLBN_CUDA_OK(cudaIpcGetMemHandle(&ipc, h->base));
  memcpy(handle, &ipc, 64);
  return 0;
}

int xlbn_halo_connect(xlbn_halo* h, const unsigned char lo_handle[64], const unsigned char hi_handle[64], int same_process) {
  XLBN_RANGE("xlbn_halo_connect");
  if (!h || !lo_handle || !hi_handle) return fail(XLBN_E_ARG, "halo_connect: NULL");
  if (same_process) {
    memcpy(&h->peer_lo, lo_handle, sizeof(char*));
    memcpy(&h->peer_hi, hi_handle, sizeof(char*));
    h->ipc_lo = h->ipc_hi = false;
  } else {
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, lo_handle, 64);
    XLBN_CUDA_OK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&h->peer_lo), ipc, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_lo = true;
    if (memcmp(lo_handle, hi_handle, 64) == 0) {  // 2-rank ring: both neighbours are the same slab
      h->peer_hi = h->peer_lo;
      h->ipc_hi = false;
    } else {
      memcpy(&ipc, hi_handle, 64);
      XLBN_CUDA_OK(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&h->peer_hi), ipc, cudaIpcMemLazyEnablePeerAccess));
      h->ipc_hi = true;
    }
  }
  h->connected = true;
  return 0;
}

int xlbn_halo_push(xlbn_halo* h, const void* f, const xlbn_domain* dom, int timestep, void* stream) {
  XLBN_RANGE("xlbn_halo_push");
  if (!h || !f || !dom) return fail(XLBN_E_ARG, "halo_push: NULL");
  if (!h->connected) return fail(XLBN_E_STATE, "halo_push: halo is not connected");
  if (int e = halo_check_alive(h, "halo_push")) return e;
  if (dom->ny != h->ny || dom->nz != h->nz || dom->nx <= 0) return fail(XLBN_E_SHAPE, "halo_push: dims do not match the halo");
  const long long plane = (long long)h->ny * h->nz;
  const long long n = plane * dom->nx;
  const int esize = (int)dtype_size(h->store_dtype);
  const int parity = timestep & 1;
  const unsigned blocks = (unsigned)((plane + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  char* to_hi = halo_ghost(h, h->peer_hi, parity, 0);  // my plane nx-1, c_x = +1  ->  hi neighbour's plane "x = -1"
  char* to_lo = halo_ghost(h, h->peer_lo, parity, 1);  // my plane 0,    c_x = -1  ->  lo neighbour's plane "x = nx"
  const char* fc = static_cast<const char*>(f);
  if (h->lattice == XLBN_D3Q19) {
    halo_push_kernel<D3Q19><<<blocks, 256, 0, st>>>(fc, to_hi, +1, dom->nx - 1, plane, n, esize);
    halo_push_kernel<D3Q19><<<blocks, 256, 0, st>>>(fc, to_lo, -1, 0, plane, n, esize);
  } else if (h->lattice == XLBN_D2Q9) {
    halo_push_kernel<D2Q9X><<<blocks, 256, 0, st>>>(fc, to_hi, +1, dom->nx - 1, plane, n, esize);
    halo_push_kernel<D2Q9X><<<blocks, 256, 0, st>>>(fc, to_lo, -1, 0, plane, n, esize);
  } else {
    halo_push_kernel<D3Q27><<<blocks, 256, 0, st>>>(fc, to_hi, +1, dom->nx - 1, plane, n, esize);
    halo_push_kernel<D3Q27><<<blocks, 256, 0, st>>>(fc, to_lo, -1, 0, plane, n, esize);
  }
  XLBN_LAUNCH_OK("halo_push_kernel");
  return 0;
}

int xlbn_halo_signal(xlbn_halo* h, int timestep, void* stream) {
  XLBN_RANGE("xlbn_halo_signal");
  if (!h) return fail(XLBN_E_ARG, "halo_signal: NULL");
  if (!h->connected) return fail(XLBN_E_STATE, "halo_signal: halo is not connected");
  if (int e = halo_check_alive(h, "halo_signal")) return e;
  // I am the hi neighbour of my lo neighbour (its flags[1]) and the lo neighbour of my hi neighbour (its flags[0])
  halo_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(halo_flags(h, h->peer_lo) + 1, halo_flags(h, h->peer_hi) + 0, timestep);
  XLBN_LAUNCH_OK("halo_signal_kernel");
  return 0;
}

int xlbn_halo_wait(xlbn_halo* h, int timestep, void* stream) {
  XLBN_RANGE("xlbn_halo_wait");
  if (!h) return fail(XLBN_E_ARG, "halo_wait: NULL");
  if (!h->connected) return fail(XLBN_E_STATE, "halo_wait: halo is not connected");
  if (int e = halo_check_alive(h, "halo_wait")) return e;
  halo_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(halo_flags(h, h->base), timestep, h->timeout_ns, h->timed_out_dev);
  XLBN_LAUNCH_OK("halo_wait_kernel");
  return 0;
}

int xlbn_halo_ghost_ptr(xlbn_halo* h, void** ptr, long long* bytes) {
  if (!h || !ptr || !bytes) return fail(XLBN_E_ARG, "halo_ghost_ptr: NULL");
  *ptr = h->base;
  *bytes = (long long)h->block_bytes;
  return 0;
}

}  // extern "C"
