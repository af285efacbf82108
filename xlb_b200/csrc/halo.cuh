// Internal definition of the x-slab halo handle (shared by halo.cu and api.cu).
#pragma once

#include "common.cuh"

struct xlbn_halo {
  int lattice, store_dtype, ny, nz, ndir;
  size_t plane_bytes;   // ny * nz * sizeof(store)
  size_t flags_offset;  // 2 parities * 2 faces * ndir planes, rounded up to 256 B
  size_t block_bytes;
  char* base;     // own ghost block (cudaMalloc)
  char* peer_lo;  // lo neighbour's ghost block, mapped into this process / device
  char* peer_hi;
  bool ipc_lo, ipc_hi, connected;
  int device;
  // Dead-neighbour detection: a device-side wait that gives up after `timeout_ns` writes the step it was waiting for into
  // *timed_out (pinned, mapped host memory), and every later xlbn_step / xlbn_halo_* call on this handle returns XLBN_E_STATE.
  volatile int* timed_out;  // host view; 0 = never
  int* timed_out_dev;       // device view of the same word
  long long timeout_ns;
};

namespace xlbn {

// ghost plane set: face 0 = plane "x = -1" (populations with c_x = +1), face 1 = plane "x = nx" (c_x = -1)
inline char* halo_ghost(const xlbn_halo* h, char* base, int parity, int face) {
  return base + ((size_t)(parity * 2 + face) * h->ndir) * h->plane_bytes;
}
// flags[0] is written by the lo neighbour, flags[1] by the hi neighbour, flags[2] = wait-timeout marker (device copy)
inline int* halo_flags(const xlbn_halo* h, char* base) { return reinterpret_cast<int*>(base + h->flags_offset); }

// < 0 (XLBN_E_STATE) once a wait on this handle has timed out: the ghosts it read were stale, results after that step are invalid
int halo_check_alive(const xlbn_halo* h, const char* where);

}  // namespace xlbn
