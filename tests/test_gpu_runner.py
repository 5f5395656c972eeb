"""GPU: the shell around the fluid step -- the headless runner `blub_run` (fast-forward protocol of
src/simulation_controller.rs:96-157), its solver-statistics history (the GUI's plots, src/gui/mod.rs:177-210, pressure_solver.rs:148-209),
its Chrome trace with the reference's profiler scope labels (src/gui/mod.rs:422-440,487-491), its particle dump, and the
renderer hand-off `blub_fluid_view` (the ten bindings of HybridFluid::bind_group_renderer, hybrid_fluid.rs:351-369,700-713)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import blub_b200
from blub_b200 import fluid as F
from tests import util
from tests.util import DT

pytestmark = pytest.mark.gpu

RUNNER = os.path.join(os.path.dirname(F.lib_path()), "blub_run")
STEPS = 32
# the scope labels below "HybridFluid step" (hybrid_fluid.rs:780-973), in submission order
SCOPES = ["transfer particle velocity to grid", "compute divergence", "primary pressure solver (divergence)", "Particle Binning",
          "make velocity grid divergence free", "extrapolate velocity grid", "clear marker & linked list grids",
          "advect particles & write new linked list grid", "density projection: set boundary marker",
          "density projection: compute density error via gather", "secondary pressure solver (density)", "compute position change",
          "extrapolate velocity grid", "correct particle density error"]


@pytest.fixture(scope="module")
def run(tmp_path_factory):
    out = tmp_path_factory.mktemp("blub_run")
    paths = {k: str(out / f"{k}.{'f32' if k == 'dump' else 'json'}") for k in ("stats", "trace", "dump")}
    assert os.path.exists(RUNNER), "blub_run is built by python -m blub_b200.build"
    cmd = [RUNNER, util.scene_path("dam_small"), "--steps", str(STEPS), "--solver", "1e-4", "128", "4", "--stats", paths["stats"], "--trace", paths["trace"],
           "--dump", paths["dump"]]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout, paths


def test_runner_follows_the_fast_forward_protocol(run):
    stdout, _ = run
    # batches of 16 steps with a wait after each (simulation_controller.rs:112,140), then the summary line (:150-156)
    assert stdout.count("simulation fast forwarding batch finished") == STEPS // 16
    assert "progress 16/32" in stdout and "progress 32/32" in stdout
    assert "Fast forward of" in stdout and "took" in stdout and "to compute" in stdout
    assert "40000 particles" in stdout and "grid 32x32x32" in stdout


def test_runner_statistics_history_matches_a_ctypes_run(run):
    _, paths = run
    stats = json.load(open(paths["stats"]))
    assert stats["steps"] == STEPS and len(stats["velocity"]) == STEPS and len(stats["density"]) == STEPS
    f = blub_b200.HybridFluid.from_scene(util.scene_path("dam_small"))
    util.tight_solver(f)
    for _ in range(STEPS):
        f.step(DT)
    f.synchronize()
    f.update_statistics()
    for name, which in (("velocity", 0), ("density", 1)):
        mine = f.pressure_solver_stats(which)
        assert len(mine) == STEPS
        e, it = f.last_solve(which)
        assert mine[-1][1] == it and abs(mine[-1][0] - e * DT) <= 1e-6 * max(1.0, e * DT)  # error = max|r| * dt (pressure_solver.rs:162)
        theirs = [(s["error"], s["iteration_count"]) for s in stats[name]]
        # converged solves stop at a multiple of 4 iterations below the 1e-4 tolerance.  Two runs of the same scene differ in the last bits
        # (float atomics of the scatters), which moves a stop decision by one check interval now and then
        assert all(i % 4 == 0 and 0 < i <= 128 for _, i in theirs)
        assert all(e_ <= 1e-4 * 1.0001 or i == 128 for e_, i in theirs)
        assert abs(theirs[0][1] - mine[0][1]) <= 4
        assert sum(abs(a[1] - b[1]) <= 4 for a, b in zip(theirs, mine)) >= STEPS - 4


def test_runner_trace_has_the_reference_scope_labels(run):
    _, paths = run
    ev = json.load(open(paths["trace"]))["traceEvents"]
    assert ev[0]["name"] == "HybridFluid step" and ev[0]["ph"] == "X"
    assert [e["name"] for e in ev[1:]] == SCOPES
    assert all(e["dur"] >= 0 for e in ev) and abs(sum(e["dur"] for e in ev[1:]) - ev[0]["dur"]) <= 1e-3 * ev[0]["dur"] + 0.5
    ts = [e["ts"] for e in ev[1:]]
    assert ts == sorted(ts)
    solver = [e["dur"] for e in ev[1:] if "pressure solver" in e["name"]]
    assert len(solver) == 2 and min(solver) > 0


def test_runner_dump_matches_a_ctypes_run(tmp_path):
    """A short run (6 steps: chaotic growth of last-bit differences stays far below the tolerance) dumped by the runner against the same
    steps driven through ctypes."""
    dump_path = str(tmp_path / "particles.f32")
    res = subprocess.run([RUNNER, util.scene_path("dam_small"), "--steps", "6", "--solver", "1e-4", "128", "4", "--dump", dump_path], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "wrote 40000 particles" in res.stdout
    dump = np.fromfile(dump_path, dtype=np.float32).reshape(-1, 4)
    assert dump.shape[0] == 40000
    f = blub_b200.HybridFluid.from_scene(util.scene_path("dam_small"))
    util.tight_solver(f)
    for _ in range(6):
        f.step(DT)
    mine = f.download_particles()
    d = np.abs(mine[:, :3] - dump[:, :3]).max(axis=1)
    assert np.isfinite(dump).all() and np.quantile(d, 0.999) <= 2e-3 and d.max() <= 5e-2, (np.quantile(d, 0.999), d.max())


def test_fluid_view_exports_the_ten_renderer_bindings():
    """blub_fluid_view: every pointer is readable device memory of the documented size and holds what the taps return."""
    cudart = C.CDLL("libcudart.so")
    cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    f = blub_b200.HybridFluid.from_scene(util.scene_path("dam_small"))
    for _ in range(3):
        f.step(DT)
    f.synchronize()
    v = f.view()
    npart, n = f.num_particles, f.n

    def read(ptr, count, dtype):
        out = np.empty(count, dtype=dtype)
        assert ptr, "NULL binding"
        assert cudart.cudaMemcpy(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes, 2) == 0  # cudaMemcpyDeviceToHost
        return out

    pairs = [("particles_position_ll", F.TAP_POS, npart * 4, np.float32), ("particles_velocity_x", F.TAP_VX, npart * 4, np.float32),
             ("particles_velocity_y", F.TAP_VY, npart * 4, np.float32), ("particles_velocity_z", F.TAP_VZ, npart * 4, np.float32)]
    for name, tap, count, dt in pairs:
        assert np.array_equal(read(getattr(v, name), count, dt), f.download_particles(tap).reshape(-1)), name
    grids = [("grid_velocity_x", F.TAP_UX, np.float32), ("grid_velocity_y", F.TAP_UY, np.float32), ("grid_velocity_z", F.TAP_UZ, np.float32),
             ("marker", F.TAP_MARKER, np.int8), ("pressure_from_velocity", F.TAP_P_VEL, np.float32), ("pressure_from_density", F.TAP_P_DEN, np.float32)]
    for name, tap, dt in grids:
        assert np.array_equal(read(getattr(v, name), n, dt), f.download_grid(tap).reshape(-1)), name
    m = read(v.marker, n, np.int8)
    assert set(np.unique(m)) <= {-1, 0, 1} and (m == 1).sum() > 1000
    pos = read(v.particles_position_ll, npart * 4, np.float32).reshape(-1, 4)[:, :3]
    assert pos.min() >= 1.0 and pos.max() <= 31.0
