"""GPU: the particle -> grid velocity transfer.  Two forms: the scatter (warp-aggregated float atomics into accumulator volumes; default,
faster) and the gather over per-step cell lists (deterministic: bit-identical output run to run).  They must agree with each other to
rounding, and both with the oracle (tests/test_gpu_parity.py::test_stagewise_parity_one_step, tests/test_transfer_kat.py)."""
import numpy as np
import pytest

import blub_b200
from blub_b200 import fluid as F
from oracle import oracle as O
from tests import util
from tests.util import DT, grid_close

pytestmark = pytest.mark.gpu

TAPS_U = (F.TAP_UX, F.TAP_UY, F.TAP_UZ)


def random_particles(n, dims, seed, lo=1.001, crowd=None):
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    pos = np.zeros((n, 4), dtype=np.float32)
    pos[:, 0] = rng.uniform(lo, nx - lo, n)
    pos[:, 1] = rng.uniform(lo, ny * 0.6, n)
    pos[:, 2] = rng.uniform(lo, nz - lo, n)
    if crowd is not None:  # a pile of particles inside one row of cells: more than a warp stages per row (GCAP)
        k, (cx, cy, cz) = crowd
        pos[:k, 0] = rng.uniform(cx, cx + 3.0, k)
        pos[:k, 1] = rng.uniform(cy, cy + 1.0, k)
        pos[:k, 2] = rng.uniform(cz, cz + 1.0, k)
    rows = [rng.normal(0, 3.0, (n, 4)).astype(np.float32) for _ in range(3)]
    return pos, rows


def p2g(dims, pos, rows, scatter=False, graph=False):
    f = blub_b200.HybridFluid(*dims, pos.shape[0])
    f.set_transfer_path(scatter)
    f.set_gravity_grid([0.0, -981.0, 0.0])
    f.set_particles(pos, *rows)
    f.step_stages(DT, 0, 1)
    return [f.download_grid(t) for t in TAPS_U], f.download_grid(F.TAP_MARKER)


@pytest.mark.parametrize("dims,n,crowd", [((64, 32, 48), 120000, None), ((24, 40, 32), 30000, None), ((64, 32, 48), 60000, (1500, (20, 5, 9)))])
def test_gather_p2g_is_bit_identical_run_to_run_and_matches_the_scatter(dims, n, crowd):
    pos, rows = random_particles(n, dims, seed=n, crowd=crowd)
    u_a, m_a = p2g(dims, pos, rows)
    u_b, m_b = p2g(dims, pos, rows)
    assert np.array_equal(m_a, m_b)
    for c in range(3):
        assert np.array_equal(u_a[c], u_b[c]), f"component {c}: gather output differs between two runs of the same input"
    u_s, m_s = p2g(dims, pos, rows, scatter=True)
    assert np.array_equal(m_a, m_s)  # marker from the cell lists == marker from the particle stores + boundary rule
    for c in range(3):
        mask = util.fluid_adjacent_faces(m_a, c)
        assert mask.sum() > 1000
        grid_close(u_s[c], u_a[c], f"gather vs scatter u[{c}]", rel=2e-5, abs_=1e-5, mask=mask)
        assert (u_a[c][~mask] == 0).all()  # faces away from the fluid are 0


def test_gather_p2g_handles_particles_on_the_clamp_planes_and_outside():
    """Positions outside the domain are clamped to [1, dim - 1] by both forms (memory safety, transfer_position)."""
    dims = (32, 32, 32)
    rng = np.random.default_rng(2)
    pos = np.zeros((4000, 4), dtype=np.float32)
    pos[:, :3] = rng.uniform(-3.0, 35.0, (4000, 3))
    pos[:50, 0] = 31.0   # exactly on the upper clamp plane
    pos[50:100, 1] = 1.0  # exactly on the lower one
    rows = [rng.normal(0, 2.0, (4000, 4)).astype(np.float32) for _ in range(3)]
    u_g, m_g = p2g(dims, pos, rows)
    u_s, m_s = p2g(dims, pos, rows, scatter=True)
    assert np.array_equal(m_g, m_s)
    for c in range(3):
        assert np.isfinite(u_g[c]).all()
        grid_close(u_s[c], u_g[c], f"u[{c}]", rel=2e-5, abs_=1e-5, mask=util.fluid_adjacent_faces(m_g, c))


def test_cell_lists_give_a_deterministic_stable_binning():
    """Stage 3 (binning) permutes by the cell lists: x-fastest cell order, and inside a cell the previous order is kept (the lists are
    canonicalised to ascending particle index) -- so the result is a pure function of the input, unlike a rank taken from atomics."""
    dims = (32, 32, 32)
    pos, _ = random_particles(50000, dims, seed=9)
    outs = []
    for _ in range(2):
        f = blub_b200.HybridFluid(*dims, pos.shape[0])
        f.set_rebin_frequency(1)
        f.set_particles(pos)
        f.step_stages(DT, 3, 4)
        outs.append(f.download_particles()[:, :3])
    assert np.array_equal(outs[0], outs[1])
    c = np.floor(pos[:, :3]).astype(np.int64)
    key = (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]
    want = pos[np.argsort(key, kind="stable"), :3]
    assert np.array_equal(outs[0], want)


def test_scatter_p2g_leaves_its_accumulators_zero():
    """The scatter form needs no memset: the finish pass re-zeroes every accumulator the particles touched.  Running the transfer again (and
    again after moving the particles) must therefore give the same result as a fresh fluid."""
    dims = (64, 32, 48)
    pos, rows = random_particles(60000, dims, seed=5, crowd=(800, (30, 4, 20)))
    pos2 = pos.copy()
    pos2[:, 0] = np.clip(pos2[:, 0] + 7.3, 1.001, dims[0] - 1.001)
    fresh = {}
    for key, p in (("a", pos), ("b", pos2)):
        fresh[key], _ = p2g(dims, p, rows, scatter=True)
    f = blub_b200.HybridFluid(*dims, pos.shape[0])
    f.set_gravity_grid([0.0, -981.0, 0.0])
    for key, p in (("a", pos), ("a", pos), ("b", pos2), ("a", pos)):
        f.set_particles(p, *rows)
        f.step_stages(DT, 0, 1)
        m = f.download_grid(F.TAP_MARKER)
        for c, t in enumerate(TAPS_U):
            grid_close(fresh[key][c], f.download_grid(t), f"repeated scatter u[{c}] ({key})", rel=2e-5, abs_=1e-5, mask=util.fluid_adjacent_faces(m, c))


def test_full_steps_are_bit_identical_run_to_run():
    """With the gather P2G nothing in the velocity half of a step depends on the arrival order of atomics; with rebinning off and the same
    input two runs of stages 0-8 must agree bit for bit."""
    res = []
    for _ in range(2):
        f = blub_b200.HybridFluid.from_scene(util.scene_path("dam_small"))
        f.set_transfer_path("gather")
        f.set_rebin_frequency(0)
        f.step_stages(DT, 0, 9)
        res.append((f.download_particles()[:, :3], f.download_grid(F.TAP_P_VEL), f.download_grid(F.TAP_MARKER)))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
