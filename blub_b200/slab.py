"""Host-side helpers of the z-slab decomposition: which planes a rank owns, how a global volume maps to
a rank's local volume with ghost planes, and the one-time exchange of the peer-visible windows between processes.
Pure NumPy / torch.distributed: the same code runs under gloo on CPU (tests) and under nccl on the GPU box (bench)."""
from __future__ import annotations

import numpy as np

HALO = 4  # SLAB_HALO in csrc/common.cuh


def owned_range(rank: int, world: int, nz_global: int):
    """Global z range [z0, z1) owned by `rank`; equal slabs, whole 4-plane tiles."""
    if nz_global % world:
        raise ValueError("nz must be divisible by the number of slabs")
    nz_owned = nz_global // world
    if nz_owned % HALO:
        raise ValueError("owned planes per slab must be a multiple of 4 (one solver tile)")
    return rank * nz_owned, (rank + 1) * nz_owned


def local_view(glob: np.ndarray, rank: int, world: int, fill=0) -> np.ndarray:
    """Global [NZ, ny, nx] volume -> the rank's local volume: HALO ghost planes, owned planes, HALO ghost planes.
    Ghost planes outside the global domain are `fill` (0 == SOLID for markers, 0.0 for fields)."""
    z0, z1 = owned_range(rank, world, glob.shape[0])
    out = np.full((z1 - z0 + 2 * HALO,) + glob.shape[1:], fill, dtype=glob.dtype)
    lo, hi = z0 - HALO, z1 + HALO
    a, b = max(lo, 0), min(hi, glob.shape[0])
    out[a - lo:b - lo] = glob[a:b]
    return out


def owned_part(local: np.ndarray) -> np.ndarray:
    return local[HALO:local.shape[0] - HALO]


def gather_global(local_owned: np.ndarray, dist) -> np.ndarray:
    """All ranks' owned planes stacked in rank order (every rank gets the global volume)."""
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local_owned)
    return np.concatenate(parts, axis=0)


def exchange_handles(my_handle: bytes, dist):
    """One 64-byte CUDA IPC handle per rank, gathered on every rank (order = rank)."""
    handles = [None] * dist.get_world_size()
    dist.all_gather_object(handles, my_handle)
    if any(h is None or len(h) != 64 for h in handles):
        raise RuntimeError("IPC handle exchange failed")
    return handles


def halo_push_reference(local: np.ndarray, rank: int, world: int, dist) -> np.ndarray:
    """What the persistent kernel's P2P pushes do, spelled out with send/recv: every rank's first/last owned plane lands
    in the neighbours' adjacent ghost plane.  Used by the CPU (gloo) tests as the specification of the exchange."""
    import torch

    out = local.copy()
    nz_owned = local.shape[0] - 2 * HALO
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(local[HALO])), rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(local[HALO + nz_owned - 1])), rank + 1))
    if rank > 0:
        buf = torch.empty(local.shape[1:], dtype=torch.from_numpy(local[:1]).dtype)
        dist.recv(buf, rank - 1)
        out[HALO - 1] = buf.numpy()
    if rank < world - 1:
        buf = torch.empty(local.shape[1:], dtype=torch.from_numpy(local[:1]).dtype)
        dist.recv(buf, rank + 1)
        out[HALO + nz_owned] = buf.numpy()
    for r in reqs:
        r.wait()
    return out
