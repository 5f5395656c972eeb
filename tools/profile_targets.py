"""Short GPU workloads to run under ncu (never a source of bench numbers).

    python tools/profile_targets.py step [scene] [steps]   # a few full steps
    python tools/profile_targets.py pcg [n] [solves] [path] # the PCG roofline microbench (all-fluid n^3 box)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blub_b200  # noqa: E402
from blub_b200 import fluid as F  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "step"
if mode == "step":
    scene = sys.argv[2] if len(sys.argv) > 2 else "dam_256"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    f = blub_b200.HybridFluid.from_scene(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".json"))
    for _ in range(steps):
        f.step(F.DT_120HZ)
    f.synchronize()
    print("launches", blub_b200.kernel_launch_count())
elif mode == "stages_sharded":
    # per-stage times of the sharded step: `world` slabs in one process (peer access), one host thread per slab
    import threading

    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    STAGES = F.STAGES
    n, cap = 256, 33000000
    for a in range(world):
        for b in range(world):
            if a != b:
                F.enable_peer_access(a, b)
    slabs = [blub_b200.HybridFluid.create_slab(n, n, n, cap, rank=k, world=world, device=k) for k in range(world)]
    wins = [s.slab_window()[0] for s in slabs]
    for s in slabs:
        s.attach_slab_peers(wins)
        s.add_fluid_cube([0.0, 0.0, 0.0], [n / 2.0, n / 4.0, float(n * world)])
        s.set_gravity_grid([0.0, -1962.0, 0.0])
    for _ in range(3):
        for s in slabs:
            s.step(F.DT_120HZ)
    for s in slabs:
        s.synchronize()
    reps, acc = 5, [np.zeros(14) for _ in slabs]
    for _ in range(reps):
        def run(k):
            acc[k] += np.array(slabs[k].step_timed(F.DT_120HZ))
        th = [threading.Thread(target=run, args=(k,)) for k in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
    for i, name in enumerate(STAGES):
        print(f"{name:24s} " + "  ".join(f"{a[i] / reps:8.3f}" for a in acc) + " ms")
    print(f"{'total':24s} " + "  ".join(f"{a.sum() / reps:8.3f}" for a in acc) + " ms (eager, per rank)")
    print("errors", [s.slab_error() for s in slabs], "particles", [s.num_particles for s in slabs])
elif mode == "cellstats":
    # python tools/profile_targets.py cellstats [scene] [steps ...]: how crowded do cells get (particles per cell) over a run
    scene = sys.argv[2] if len(sys.argv) > 2 else "dam_256"
    checkpoints = [int(a) for a in sys.argv[3:]] or [3, 56, 110]
    f = blub_b200.HybridFluid.from_scene(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".json"))
    done = 0
    for cp in checkpoints:
        while done < cp:
            f.step(F.DT_120HZ)
            done += 1
        p = f.download_particles()[:, :3]
        c = np.floor(p).astype(np.int64)
        cnt = np.bincount((c[:, 2] * f.ny + c[:, 1]) * f.nx + c[:, 0], minlength=f.n)
        nz = cnt[cnt > 0]
        rows = cnt.reshape(f.nz, f.ny, f.nx)
        seg = rows.reshape(f.nz, f.ny, f.nx // 32, 32).sum(-1) if f.nx % 32 == 0 else rows.sum(-1, keepdims=True)
        print(f"step {cp}: fluid cells {nz.size} ({100.0 * nz.size / f.n:.1f} %), particles per fluid cell mean {nz.mean():.2f} q99.9 {np.quantile(nz, 0.999):.0f} max {cnt.max()}, "
              f"cells > 32: {(cnt > 32).sum()}, > 100: {(cnt > 100).sum()}, > 1000: {(cnt > 1000).sum()}; particles per 32-cell row segment max {seg.max()}, > 384: {(seg > 384).sum()}", flush=True)
elif mode == "pcg_overhead":
    # the fixed cost of an iteration: a 256^3 grid whose only fluid is one small block (one tile / a few columns): ms per solve / 66 = barrier +
    # reduction cost of one phase
    n = 256
    for blob in (8, 64):
        f = blub_b200.HybridFluid(n, n, n, 8)
        m = np.full((n, n, n), -1, dtype=np.int8)
        m[0], m[-1], m[:, 0], m[:, -1], m[:, :, 0], m[:, :, -1] = 0, 0, 0, 0, 0, 0
        m[100:100 + blob, 100:100 + blob, 128:128 + blob] = 1
        b = np.random.default_rng(1).uniform(-1, 1, (n, n, n)).astype(np.float32)
        f.upload_grid(F.TAP_MARKER, m)
        f.upload_grid(F.TAP_RESIDUAL, b)
        f.set_solver_config(0, 0.0, 32, 4)
        ms = sorted(f.time_solve(0, F.DT_120HZ, 7))
        print(f"fluid block {blob}^3: ms per solve {ms[len(ms) // 2]:.4f} = {1e3 * ms[len(ms) // 2] / 66:.2f} us per phase; work {f.solver_work()}", flush=True)
        f.close()
elif mode == "stages":
    # python tools/profile_targets.py stages [scene] [checkpoint steps ...]: stage times after so many steps (default 3)
    scene = sys.argv[2] if len(sys.argv) > 2 else "dam_256"
    checkpoints = [int(a) for a in sys.argv[3:]] or [3]
    f = blub_b200.HybridFluid.from_scene(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".json"))
    if os.environ.get("BLUB_REBIN"):  # experiment: the reference's particle_rebinning_step_frequency (default 60)
        f.set_rebin_frequency(int(os.environ["BLUB_REBIN"]))
    STAGES = F.STAGES
    done, cols, reps = 0, [], 3
    for cp in checkpoints:
        while done < cp:
            f.step(F.DT_120HZ)
            done += 1
        acc = np.zeros(14)
        for _ in range(reps):
            acc += np.array(f.step_timed(F.DT_120HZ))
            done += 1
        cols.append(acc / reps)
        print(f"# after {cp} steps: solver work {f.solver_work()}", flush=True)
    print(f"{'after steps':24s} " + " ".join(f"{c:8d}" for c in checkpoints))
    for i, name in enumerate(STAGES):
        print(f"{name:24s} " + " ".join(f"{c[i]:8.3f}" for c in cols) + " ms")
    print(f"{'total':24s} " + " ".join(f"{c.sum():8.3f}" for c in cols) + " ms (eager launches, events between stages)")
    f.update_statistics()
    print("solver stats", f.pressure_solver_stats(0)[-1], f.pressure_solver_stats(1)[-1], "solver work", f.solver_work())
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    solves = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    f = blub_b200.HybridFluid(n, n, n, 8)
    m = np.zeros((n, n, n), dtype=np.int8)
    m[1:-1, 1:-1, 1:-1] = 1
    b = np.random.default_rng(1234).uniform(-1, 1, (n, n, n)).astype(np.float32)
    b -= b[m == 1].mean(dtype=np.float64).astype(np.float32)
    b[m != 1] = 0
    f.upload_grid(F.TAP_MARKER, m)
    f.upload_grid(F.TAP_RESIDUAL, b)
    if len(sys.argv) > 4:
        f.set_solver_path(int(sys.argv[4]) if sys.argv[4].isdigit() else sys.argv[4])
    f.set_solver_config(0, 0.0, 32, 4)
    print("ms per solve", f.time_solve(0, F.DT_120HZ, solves))
