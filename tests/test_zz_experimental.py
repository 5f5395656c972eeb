"""Opt-in kernel variants that are NOT the default and have not been measured yet.  They only run with BLUB_EXPERIMENTAL=1
(next round's first GPU session); each must reproduce the default kernel bit for bit before it may be timed."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import DT

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("BLUB_EXPERIMENTAL"), reason="set BLUB_EXPERIMENTAL=1 to run the experimental variants")]


def test_byte_mask_extrapolation_is_bit_identical():
    import blub_b200
    from blub_b200 import fluid as F

    nx, ny, nz = 64, 40, 48
    rng = np.random.default_rng(3)
    m = np.full((nz, ny, nx), O.AIR, dtype=np.int8)
    m[rng.random((nz, ny, nx)) < 0.02] = O.FLUID                 # spray
    m[8:30, 4:20, 10:50][rng.random((22, 16, 40)) < 0.9] = O.FLUID  # a ragged body
    m[rng.random((nz, ny, nx)) < 0.01] = O.SOLID
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    u = [rng.uniform(-5, 5, (nz, ny, nx)).astype(np.float32) for _ in range(3)]
    out = {}
    for mode in ("default", "bytes"):
        if mode == "bytes":
            os.environ["BLUB_EXTRAPOLATE"] = "bytes"
        else:
            os.environ.pop("BLUB_EXTRAPOLATE", None)
        try:
            f = blub_b200.HybridFluid(nx, ny, nz, 8)
        finally:
            os.environ.pop("BLUB_EXTRAPOLATE", None)
        f.upload_grid(F.TAP_MARKER, m)
        for c, t in enumerate((F.TAP_UX, F.TAP_UY, F.TAP_UZ)):
            f.upload_grid(t, u[c])
        f.step_stages(DT, 8, 9)  # boundary marker: rebuilds the occupancy maps (and the face-validity bytes)
        f.step_stages(DT, 5, 6)
        out[mode] = [f.download_grid(t) for t in (F.TAP_UX, F.TAP_UY, F.TAP_UZ)]
    changed = 0
    for c in range(3):
        assert np.array_equal(out["default"][c], out["bytes"][c])
        changed += int((out["default"][c] != u[c]).sum())
    assert changed > 1000


def test_warp_aggregated_scatters_match_the_default():
    """BLUB_SCATTER=aggregate: same (face, particle) pairs and weights, summed per run of equal dual cells before the reductions."""
    import blub_b200
    from blub_b200 import fluid as F
    from tests import util
    from tests.util import grid_close

    results = {}
    rng = np.random.default_rng(11)
    rows = None
    for mode in ("default", "aggregate"):
        os.environ.pop("BLUB_SCATTER", None)
        f = blub_b200.HybridFluid.from_scene(util.scene_path("dam_small"))
        f.set_rebin_frequency(0)
        f.set_graph_replay(False)
        pos = f.download_particles().copy()
        if rows is None:
            rows = [rng.normal(0, 3.0, pos.shape).astype(np.float32) for _ in range(3)]
        f.set_particles(pos, *rows)
        if mode == "aggregate":
            os.environ["BLUB_SCATTER"] = "aggregate"
        try:
            f.step_stages(DT, 0, 1)
            u = [f.download_grid(t) for t in (F.TAP_UX, F.TAP_UY, F.TAP_UZ)]
            m = f.download_grid(F.TAP_MARKER)
            f.step_stages(DT, 9, 10)
            rhs = f.download_grid(F.TAP_RESIDUAL)
        finally:
            os.environ.pop("BLUB_SCATTER", None)
        results[mode] = (u, m, rhs)
    assert np.array_equal(results["default"][1], results["aggregate"][1])
    fl = results["default"][1] == O.FLUID
    for c in range(3):
        grid_close(results["default"][0][c], results["aggregate"][0][c], f"P2G u[{c}]", rel=1e-5, abs_=1e-5, mask=util.fluid_adjacent_faces(results["default"][1], c))
    grid_close(results["default"][2], results["aggregate"][2], "density rhs", rel=1e-5, abs_=1e-3, mask=fl)


@pytest.mark.parametrize("nx,fill", [(64, 0.5), (128, 0.9), (96, 0.05)])
def test_brick_granular_pcg_matches_the_tile_kernel(nx, fill):
    """Solver path 5 (one warp per 32x4x4 brick): same bodies, same barriers; only the work unit and with it the grouping of the
    per-block partial sums differ, so the iterates agree to rounding."""
    import blub_b200
    from blub_b200 import fluid as F
    from tests.util import grid_close

    ny, nz = 40, 24
    rng = np.random.default_rng(nx)
    m = np.full((nz, ny, nx), O.AIR, dtype=np.int8)
    m[rng.random((nz, ny, nx)) < fill] = O.FLUID
    m[rng.random((nz, ny, nx)) < 0.03] = O.SOLID
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    b = rng.uniform(-1, 1, (nz, ny, nx)).astype(np.float32)
    out = {}
    for path in (True, "brick"):
        f = blub_b200.HybridFluid(nx, ny, nz, 8)
        f.set_solver_path(path)
        f.set_solver_config(0, 0.0, 24, 4)
        f.upload_grid(F.TAP_MARKER, m)
        for rep in range(2):  # second solve: warm start
            f.upload_grid(F.TAP_RESIDUAL, b)
            f.solve_only(0, DT)
        out[path] = (f.download_grid(F.TAP_P_VEL), f.last_solve(0))
    assert out[True][1][1] == out["brick"][1][1] == 24
    grid_close(out[True][0], out["brick"][0], "pressure brick vs tile", rel=2e-3, abs_=1e-4)
    assert (out["brick"][0][m != O.FLUID] == 0).all()
