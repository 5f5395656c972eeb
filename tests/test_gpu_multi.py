"""Multi-GPU: the z-slab sharded PCG (P2P ghost-plane pushes + mailbox all-reduce inside the persistent kernel)
must reproduce the single-GPU solve of the same global problem.  Needs >= 2 GPUs (skipped otherwise)."""
import numpy as np
import pytest

import blub_b200
from blub_b200 import fluid as F
from blub_b200 import slab
from oracle import oracle as O
from tests.util import DT, grid_close

pytestmark = pytest.mark.gpu
HALO = slab.HALO


def _gpu_count():
    import torch
    return torch.cuda.device_count()


def make_slabs(world, nx, ny, nz_owned):
    for a in range(world):
        for b in range(world):
            if a != b:
                F.enable_peer_access(a, b)
    slabs = [blub_b200.HybridFluid.create_slab(nx, ny, nz_owned, 8, rank=k, world=world, device=k) for k in range(world)]
    windows = [s.slab_window()[0] for s in slabs]
    for s in slabs:
        s.attach_slab_peers(windows)
    return slabs


def local_view(glob, k, nz_owned, fill=0):
    return slab.local_view(glob, k, glob.shape[0] // nz_owned, fill)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_pcg_matches_single_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    nx, ny, nz_owned = 64, 48, 16
    NZ = world * nz_owned
    rng = np.random.default_rng(world)
    m = np.full((NZ, ny, nx), O.AIR, dtype=np.int8)
    m[rng.random((NZ, ny, nx)) < 0.75] = O.FLUID
    m[rng.random((NZ, ny, nx)) < 0.03] = O.SOLID
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    b = rng.uniform(-1, 1, (NZ, ny, nx)).astype(np.float32)

    ref = blub_b200.HybridFluid(nx, ny, NZ, 8, device=0)
    ref.set_solver_config(0, 1e-3, 64, 4)
    ref.upload_grid(F.TAP_MARKER, m)
    slabs = make_slabs(world, nx, ny, nz_owned)
    for k, s in enumerate(slabs):
        s.set_solver_config(0, 1e-3, 64, 4)
        s.upload_grid(F.TAP_MARKER, local_view(m, k, nz_owned))
    for rep in range(2):  # second solve: warm start through the pushed ghost planes of p
        ref.upload_grid(F.TAP_RESIDUAL, b)
        ref.solve_only(0, DT)
        for k, s in enumerate(slabs):
            s.upload_grid(F.TAP_RESIDUAL, local_view(b, k, nz_owned))
        for s in slabs:  # all ranks enqueue, then all wait: the kernels talk to each other
            s.solve_only(0, DT)
        e_ref, it_ref = ref.last_solve(0)
        stats = [s.last_solve(0) for s in slabs]
        assert all(it == stats[0][1] for _, it in stats), stats
        assert abs(stats[0][1] - it_ref) <= 4, (stats, it_ref)
        assert all(abs(e - stats[0][0]) <= 1e-6 * max(1.0, stats[0][0]) for e, _ in stats), stats  # same scalars on every rank
        p_ref = ref.download_grid(F.TAP_P_VEL)
        for k, s in enumerate(slabs):
            p = s.download_grid(F.TAP_P_VEL)
            own = p[HALO:HALO + nz_owned]
            grid_close(p_ref[k * nz_owned:(k + 1) * nz_owned], own, f"rank {k} pressure (solve {rep})", rel=3e-3, abs_=1e-4)
            # ghost planes hold the neighbours' boundary planes
            if k > 0:
                nb = slabs[k - 1].download_grid(F.TAP_P_VEL)
                assert np.array_equal(p[HALO - 1], nb[HALO + nz_owned - 1])
            if k < world - 1:
                nb = slabs[k + 1].download_grid(F.TAP_P_VEL)
                assert np.array_equal(p[HALO + nz_owned], nb[HALO])
