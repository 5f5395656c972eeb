import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blub_b200
from blub_b200 import fluid as F
nx, ny, nz = 128, 40, 24
rng = np.random.default_rng(21)
m = np.full((nz, ny, nx), -1, dtype=np.int8)
m[rng.random((nz, ny, nx)) < 0.8] = 1
m[rng.random((nz, ny, nx)) < 0.04] = 0
m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
b = rng.uniform(-1, 1, (nz, ny, nx)).astype(np.float32)
for max_it in [0, 1, 2, 3, 4, 9]:
    res = {}
    for path in ("tma", True):
        g = blub_b200.HybridFluid(nx, ny, nz, 8)
        g.set_solver_path(path)
        g.set_solver_config(0, 0.0, max_it, 4)
        g.upload_grid(F.TAP_MARKER, m)
        g.upload_grid(F.TAP_RESIDUAL, b)
        g.solve_only(0, F.DT_120HZ)
        res[path] = (g.download_grid(F.TAP_P_VEL), g.download_grid(F.TAP_RESIDUAL), g.last_solve(0))
    dp = np.abs(res["tma"][0] - res[True][0])
    dr = np.abs(res["tma"][1] - res[True][1])
    z, y, x = np.unravel_index(np.argmax(dp), dp.shape)
    print(f"max_it={max_it}: max|dp|={dp.max():.3e} at (x={x},y={y},z={z}) scale {np.abs(res[True][0]).max():.3e}; max|dr|={dr.max():.3e}; stats {res['tma'][2]} vs {res[True][2]}")
    if dp.max() > 1e-4:
        bad = np.argwhere(dp > 0.1 * dp.max())
        print("   bad cells: n=%d  x range %d..%d  y values %s  z values %s" % (len(bad), bad[:, 2].min(), bad[:, 2].max(), sorted(set(bad[:, 1]))[:12], sorted(set(bad[:, 0]))[:12]))
