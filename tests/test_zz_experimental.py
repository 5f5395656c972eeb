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
