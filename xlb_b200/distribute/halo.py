"""Halo exchange of the x-slab decomposition.

Two implementations of the same ring exchange (reference: xlb/distribute/distribute.py:23-44; which populations cross a
face: velocity_set.right_indices (c_x = +1) leave through the high-x face, left_indices (c_x = -1) through the low-x
face; ring with periodic wrap rank n-1 <-> 0):

* ``PeerHalo`` — the product path on GPUs.  Each slab owns a compact ghost block (cudaMalloc, exported with CUDA IPC and
  mapped by both neighbours).  The step kernel of a slab's two face planes stores the outgoing populations straight into
  the neighbours' ghost planes over NVLink (peer stores inside the compute kernel), a one-thread kernel publishes a step
  counter, and the next step's face kernels spin on the counters on the device.  The interior planes run concurrently on
  the caller's stream; no host synchronisation, no NCCL call and no pack/unpack pass on the data path.
* ``exchange_wrapped_faces`` / ``exchange_ghost_planes`` — point-to-point `torch.distributed` versions (NCCL or gloo).
  Used by `distribute(Stream ...)`, by the CPU (gloo) tests of the host-side logic, and as a cross-check of PeerHalo.
"""

import ctypes as C

import torch
import torch.distributed as dist

from xlb_b200 import native


def ring_neighbours(rank: int, world: int):
    """(lo, hi) ranks of a slab in the periodic ring."""
    return (rank - 1) % world, (rank + 1) % world


def _p2p(send_hi, send_lo, rank, world):
    """Send `send_hi` to rank+1 and `send_lo` to rank-1; return (from_lo, from_hi)."""
    lo, hi = ring_neighbours(rank, world)
    from_lo, from_hi = torch.empty_like(send_hi), torch.empty_like(send_lo)
    ops = [
        dist.P2POp(dist.isend, send_hi, hi),
        dist.P2POp(dist.isend, send_lo, lo),
        dist.P2POp(dist.irecv, from_lo, lo),
        dist.P2POp(dist.irecv, from_hi, hi),
    ]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return from_lo, from_hi


def exchange_wrapped_faces(result, velocity_set, rank, world):
    """Reference post-streaming fix-up (distribute.py:26-44): after a LOCAL periodic roll, plane 0 of the c_x = +1
    populations and plane -1 of the c_x = -1 populations hold values that belong to the ring neighbours; swap them."""
    right = torch.as_tensor(velocity_set.right_indices, device=result.device)
    left = torch.as_tensor(velocity_set.left_indices, device=result.device)
    send_hi = result[right, :1].contiguous()  # belongs to plane 0 of rank+1
    send_lo = result[left, -1:].contiguous()  # belongs to plane -1 of rank-1
    from_lo, from_hi = _p2p(send_hi, send_lo, rank, world)
    result[right, :1] = from_lo
    result[left, -1:] = from_hi
    return result


def exchange_ghost_planes(f, velocity_set, rank, world):
    """Pre-streaming form used by the fused step: returns (ghost_lo, ghost_hi) = the neighbours' face populations that a
    pull across the slab faces reads: ghost_lo = plane "x = -1" (c_x = +1 populations of rank-1's last plane),
    ghost_hi = plane "x = nx" (c_x = -1 populations of rank+1's first plane)."""
    right = torch.as_tensor(velocity_set.right_indices, device=f.device)
    left = torch.as_tensor(velocity_set.left_indices, device=f.device)
    send_hi = f[right, -1].contiguous()
    send_lo = f[left, 0].contiguous()
    return _p2p(send_hi, send_lo, rank, world)


class PeerHalo:
    """Device-resident ghost planes of one slab, connected to the ring neighbours through CUDA IPC."""

    def __init__(self, grid, velocity_set, precision_policy, dims, group=None):
        self.rank, self.world = grid.rank, grid.nDevices
        self.group = group
        self.device = grid.device
        nx, ny, nz = dims
        self.handle = C.c_void_p()
        L = native.lib()
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)  # the ghost block and the IPC mappings belong to the slab's device
        native.check(L.xlbn_halo_create(velocity_set.lattice_code, precision_policy.store_precision.code, ny, nz, C.byref(self.handle)))
        mine = C.create_string_buffer(64)
        native.check(L.xlbn_halo_export(self.handle, mine))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(mine.raw), group=group)
        lo, hi = ring_neighbours(self.rank, self.world)
        native.check(L.xlbn_halo_connect(self.handle, handles[lo], handles[hi], 0))
        self.side = torch.cuda.Stream(device=self.device, priority=-1)
        self.ev_prev = torch.cuda.Event()
        self.ev_face = torch.cuda.Event()
        self.primed = False
        dist.barrier(group=group)  # every rank has mapped its neighbours before anyone stores into them

    def __del__(self):
        try:
            native.lib().xlbn_halo_destroy(self.handle)
        except Exception:
            pass

    def timed_out(self) -> bool:
        """True once a device-side wait for the neighbours' step counters gave up (default 120 s, `set_timeout` / environment
        XLBN_HALO_TIMEOUT_S).  The wait kernel never hangs the GPU: it marks the handle in mapped host memory, and from then on
        every `step` raises (XLBN_E_STATE): the step that timed out read stale ghosts.  No device synchronisation."""
        return native.lib().xlbn_halo_timed_out(self.handle) == 1

    def set_timeout(self, seconds: float):
        native.check(native.lib().xlbn_halo_set_timeout(self.handle, float(seconds)))

    def step(self, stepper_handle, f_0, f_1, bc_mask, bits, dims, omega, t):
        L = native.lib()
        nx, ny, nz = dims
        main = torch.cuda.current_stream(self.device)
        main_p, side_p = C.c_void_p(main.cuda_stream), C.c_void_p(self.side.cuda_stream)
        full = native.Domain(nx, ny, nz, 0, nx)
        if not self.primed:
            # ghosts of BOTH parities start from the initial state (solid cells are never refreshed, like f_1 = copy(f_0))
            native.check(L.xlbn_halo_push(self.handle, native.ptr(f_0), C.byref(full), t + 1, main_p))
            native.check(L.xlbn_halo_push(self.handle, native.ptr(f_0), C.byref(full), t, main_p))
            native.check(L.xlbn_halo_signal(self.handle, t, main_p))
            self.primed = True
        args = (stepper_handle, native.ptr(f_0), native.ptr(f_1), native.ptr(bc_mask), native.ptr(bits))
        if omega != getattr(self, "_omega", None):  # one step = launches on two streams: publish omega on the one both are ordered after
            native.check(L.xlbn_stepper_prepare(stepper_handle, omega, main_p))
            self._omega = omega
        if nx < 3:  # nothing to overlap
            native.check(L.xlbn_halo_wait(self.handle, t, main_p))
            native.check(L.xlbn_step(*args, C.byref(full), omega, t, self.handle, main_p))
            native.check(L.xlbn_halo_signal(self.handle, t + 1, main_p))
            return
        # face planes (need the neighbours' data, produce the neighbours' data) on the high-priority side stream ...
        self.ev_prev.record(main)
        self.side.wait_event(self.ev_prev)
        native.check(L.xlbn_halo_wait(self.handle, t, side_p))
        for x in (0, nx - 1):
            dom = native.Domain(nx, ny, nz, x, 1)
            native.check(L.xlbn_step(*args, C.byref(dom), omega, t, self.handle, side_p))
        native.check(L.xlbn_halo_signal(self.handle, t + 1, side_p))
        self.ev_face.record(self.side)
        # ... overlapped with the interior on the caller's stream
        dom = native.Domain(nx, ny, nz, 1, nx - 2)
        native.check(L.xlbn_step(*args, C.byref(dom), omega, t, None, main_p))
        main.wait_event(self.ev_face)
