"""Small workloads for compute-sanitizer (tools/sanitize.sh): every default-path kernel runs at least once on a 32^3..64^3 grid.

    python tools/sanitize_targets.py pcg       # column solver, tile kernel (sparse + dense bodies), three-kernel path, TMA path
    python tools/sanitize_targets.py step      # three full steps of the 32^3 dam break (gather P2G, cell lists, binning, all grid passes), then the scatter form
    python tools/sanitize_targets.py slab      # 2 z-slabs in one process: sharded solve + three sharded steps with migration (needs 2 GPUs)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blub_b200  # noqa: E402
from blub_b200 import fluid as F  # noqa: E402

DT = F.DT_120HZ
mode = sys.argv[1] if len(sys.argv) > 1 else "step"


def blob(nx, ny, nz, fill, seed):
    rng = np.random.default_rng(seed)
    m = np.full((nz, ny, nx), -1, dtype=np.int8)
    m[rng.random((nz, ny, nx)) < fill] = 1
    m[rng.random((nz, ny, nx)) < 0.03] = 0
    m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
    return m, rng.uniform(-1, 1, (nz, ny, nx)).astype(np.float32)


if mode == "pcg":
    for (nx, ny, nz), fill, paths in [((64, 40, 24), 0.3, (1, 6, 4, 0)), ((128, 24, 16), 0.9, (1, 2, 6)), ((24, 40, 32), 0.6, (1, 0))]:
        m, b = blob(nx, ny, nz, fill, nx)
        for path in paths:
            f = blub_b200.HybridFluid(nx, ny, nz, 8)
            f.set_solver_path(path)
            f.set_solver_config(0, 1e-3, 24, 4)
            f.upload_grid(F.TAP_MARKER, m)
            for _ in range(2):
                f.upload_grid(F.TAP_RESIDUAL, b)
                f.solve_only(0, DT)
            print("pcg", (nx, ny, nz), "path", path, f.last_solve(0), flush=True)
            f.close()
elif mode == "step":
    scene = os.path.join(ROOT, "tests", "golden", "scenes", "dam_small.json")
    for scatter in (False, True):
        f = blub_b200.HybridFluid.from_scene(scene)
        f.set_transfer_path(scatter)
        f.set_rebin_frequency(2)
        for _ in range(3):
            f.step(DT)
        f.synchronize()
        p = f.download_particles()
        print("step scatter" if scatter else "step gather", p.shape, bool(np.isfinite(p).all()), f.last_solve(0), f.last_solve(1), flush=True)
        f.close()
elif mode == "slab":
    world, nx, ny, nz_owned = 2, 64, 32, 16
    for a in range(world):
        for b in range(world):
            if a != b:
                F.enable_peer_access(a, b)
    cap = 8 * 31 * 20 * 31 + 1000
    slabs = [blub_b200.HybridFluid.create_slab(nx, ny, nz_owned, cap, rank=k, world=world, device=k) for k in range(world)]
    wins = [s.slab_window()[0] for s in slabs]
    for s in slabs:
        s.attach_slab_peers(wins)
        s.add_fluid_cube([0.0, 0.0, 0.0], [32.0, 21.0, float(world * nz_owned)])
        s.set_gravity_grid([0.0, -981.0, 0.0])
    for _ in range(3):
        for s in slabs:
            s.step(DT)
        for s in slabs:
            s.synchronize()
    print("slab", [s.num_particles for s in slabs], [s.slab_error() for s in slabs], [s.last_solve(0) for s in slabs], flush=True)
