"""Host-side logic of the x-slab decomposition on CPU: two processes, gloo backend.
 * exchange_wrapped_faces == the reference's ppermute fix-up (distribute.py:23-44): local roll + exchange == global roll
 * exchange_ghost_planes delivers exactly the planes a pull across the slab faces reads
 * the slab-aware grid hands each rank its planes and global start index."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lattice, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import xlb_b200 as xlb
        from oracle import lbm_numpy as O
        from xlb_b200.compute_backend import ComputeBackend
        from xlb_b200.distribute.halo import exchange_ghost_planes, exchange_wrapped_faces, ring_neighbours
        from xlb_b200.grid import grid_factory

        pp = xlb.PrecisionPolicy.FP32FP32
        vs = getattr(xlb.velocity_set, lattice)(precision_policy=pp, compute_backend=ComputeBackend.JAX)
        xlb.init(velocity_set=vs, default_backend=ComputeBackend.JAX, default_precision_policy=pp)
        lat = O.Lattice(lattice)
        shape = (4 * world, 5, 6)
        f = np.random.default_rng(7).random((lat.q,) + shape).astype(np.float32)  # same on every rank
        grid = grid_factory(shape, device="cpu")
        assert grid.nDevices == world and grid.rank == rank
        assert grid.local_shape == (4, 5, 6) and grid.start_index == (4 * rank, 0, 0)
        assert tuple(grid.create_field(lat.q).shape) == (lat.q, 4, 5, 6)
        x0 = grid.start_index[0]
        local = torch.as_tensor(f[:, x0 : x0 + 4].copy())

        # (1) reference semantics: local periodic roll, then swap the wrapped face planes
        rolled = torch.as_tensor(O.stream(local.numpy(), lat))
        fixed = exchange_wrapped_faces(rolled, vs, rank, world).numpy()
        ok1 = np.array_equal(fixed, O.stream(f, lat)[:, x0 : x0 + 4])

        # (2) pre-exchange form used by the fused kernel
        ghost_lo, ghost_hi = exchange_ghost_planes(local, vs, rank, world)
        lo, hi = ring_neighbours(rank, world)
        ok2 = np.array_equal(ghost_lo.numpy(), f[lat.right, (x0 - 1) % shape[0]]) and np.array_equal(ghost_hi.numpy(), f[lat.left, (x0 + 4) % shape[0]])
        out[rank] = (ok1, ok2, lo == (rank - 1) % world and hi == (rank + 1) % world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lattice", ["D3Q19", "D3Q27"])
def test_ring_exchange_world_size_2(lattice):
    world = 2
    with mp.get_context("spawn").Manager() as manager:  # never fork a multi-threaded (torch) process
        out = manager.dict()
        mp.spawn(_worker, args=(world, _free_port(), lattice, out), nprocs=world, join=True)
        assert all(all(out[r]) for r in range(world)), dict(out)
