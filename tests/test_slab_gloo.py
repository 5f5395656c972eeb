"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU path: slab partition, local views with ghost
planes, the handle exchange, and the ghost-plane push specification.  The CUDA side of the same path is covered by
tests/test_gpu_multi.py on a multi-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from blub_b200 import slab


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(99)  # same global volume on every rank
        glob = rng.standard_normal((16, 6, 8)).astype(np.float32)
        z0, z1 = slab.owned_range(rank, world, glob.shape[0])
        loc = slab.local_view(glob, rank, world)
        assert loc.shape == (z1 - z0 + 2 * slab.HALO, 6, 8)
        assert np.array_equal(slab.owned_part(loc), glob[z0:z1])
        # ghost planes beyond the domain are SOLID / zero
        if rank == 0:
            assert (loc[:slab.HALO] == 0).all()
        if rank == world - 1:
            assert (loc[-slab.HALO:] == 0).all()
        # every rank modifies its owned planes; the push must refresh exactly the adjacent ghost planes
        mine = loc.copy()
        mine[slab.HALO:-slab.HALO] += 100.0 * (rank + 1)
        pushed = slab.halo_push_reference(mine, rank, world, dist)
        full = slab.gather_global(slab.owned_part(mine), dist)
        lo, hi = z0 - 1, z1
        if rank > 0:
            assert np.array_equal(pushed[slab.HALO - 1], full[lo])
        if rank < world - 1:
            assert np.array_equal(pushed[-slab.HALO], full[hi])
        assert np.array_equal(slab.owned_part(pushed), slab.owned_part(mine))
        # IPC handle exchange: rank order, 64 bytes each
        handles = slab.exchange_handles(bytes([rank]) * 64, dist)
        assert [h[0] for h in handles] == list(range(world))
        # scalar all-reduce in rank order (what the mailbox all-reduce computes)
        vals = [None] * world
        dist.all_gather_object(vals, float(rank + 1) * 0.1)
        assert abs(sum(vals) - 0.1 * world * (world + 1) / 2) < 1e-12
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_slab_partition_and_push_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_partition_rejects_bad_shapes():
    with pytest.raises(ValueError):
        slab.owned_range(0, 3, 16)
    with pytest.raises(ValueError):
        slab.owned_range(0, 2, 12)
    assert slab.owned_range(1, 2, 16) == (8, 16)
