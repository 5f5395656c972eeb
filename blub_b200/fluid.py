"""ctypes mirror of the reference's ``HybridFluid`` (src/simulation/hybrid_fluid.rs) over libblubcore.so.

Method names, argument meaning and defaults follow the Rust interface so that parity tests read like the reference's
call sites (src/scene/mod.rs:109-144,166-214).  All compute happens in the CUDA library; nothing here falls back to
NumPy or to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# Duration::from_nanos(1e9 / 120).as_secs_f32() (simulation_controller.rs:33-39)
DT_120HZ = float(np.float32(8333333e-9))

# the 14 stages of one step, in the order of HybridFluid::step (hybrid_fluid.rs:770-977); indices for step_stages / step_timed
STAGES = [
    "p2g", "divergence_compute", "solve_velocity", "binning", "divergence_remove", "extrapolate",
    "transfer_clear", "advect", "set_boundary_marker", "density_gather_error", "solve_density",
    "position_change", "extrapolate2", "correct_particles",
]
TAP_POS, TAP_VX, TAP_VY, TAP_VZ, TAP_UX, TAP_UY, TAP_UZ, TAP_MARKER, TAP_P_VEL, TAP_P_DEN, TAP_RESIDUAL = range(11)

BLUB_OK = 0
BLUB_WARN_TRUNCATED = 100


class BlubError(RuntimeError):
    pass


class SolverConfig(C.Structure):
    """pressure_solver.rs:57-62"""

    _fields_ = [("error_tolerance", C.c_float), ("max_num_iterations", C.c_int32), ("error_check_frequency", C.c_int32)]


class SolverSample(C.Structure):
    _fields_ = [("error", C.c_float), ("iteration_count", C.c_int32)]


class Quirks(C.Structure):
    _fields_ = [("precond_mode", C.c_int32), ("reserved", C.c_int32 * 7)]


class FluidView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "particles_position_ll", "particles_velocity_x", "particles_velocity_y", "particles_velocity_z",
        "grid_velocity_x", "grid_velocity_y", "grid_velocity_z", "marker", "pressure_from_velocity", "pressure_from_density")]


class RigidObject(C.Structure):
    """BlubRigidObject: StaticObjectConfig + RigidAnimation (src/scene/models.rs:11-46) for an analytic box / sphere."""

    _fields_ = [("world_position", C.c_float * 3), ("scale", C.c_float), ("rotation_angles_deg", C.c_float * 3), ("shape", C.c_int32),
                ("half_extent", C.c_float * 3), ("has_translation", C.c_int32), ("translation_target", C.c_float * 3),
                ("translation_curve", C.c_int32), ("translation_duration", C.c_float), ("has_rotation", C.c_int32),
                ("rotation_axis", C.c_float * 3), ("rotation_deg_per_sec", C.c_float)]

    @classmethod
    def from_dict(cls, obj):
        o = cls()
        o.world_position[:] = [float(v) for v in obj["world_position"]]
        o.scale = float(obj.get("scale", 1.0))
        o.rotation_angles_deg[:] = [float(v) for v in obj.get("rotation_angles", (0, 0, 0))]
        o.shape = {"box": 0, "sphere": 1, "mesh": 2}[obj.get("shape", "box")]
        o.half_extent[:] = [float(v) for v in obj.get("half_extent", (0, 0, 0))]
        tr, rot = obj.get("translation"), obj.get("rotation")
        if tr:
            o.has_translation = 1
            o.translation_target[:] = [float(v) for v in tr["target"]]
            o.translation_curve = 1 if tr.get("curve", "Linear") == "SmoothStep" else 0
            o.translation_duration = float(tr["duration"])
        if rot:
            o.has_rotation = 1
            o.rotation_axis[:] = [float(v) for v in rot["axis"]]
            o.rotation_deg_per_sec = float(rot["deg_per_sec"])
        return o


class RigidState(C.Structure):
    _fields_ = [("centre_voxel", C.c_float * 3), ("velocity_voxel", C.c_float * 3), ("axis_scaled", C.c_float * 3), ("rotation", C.c_float * 9)]


class SceneInfo(C.Structure):
    _fields_ = [("grid_dimension", C.c_uint32 * 3), ("max_num_particles", C.c_uint32), ("grid_to_world_scale", C.c_float),
                ("world_position", C.c_float * 3), ("gravity", C.c_float * 3), ("num_fluid_cubes", C.c_uint32),
                ("num_static_objects", C.c_uint32)]


def lib_path() -> str:
    return os.path.join(_HERE, "libblubcore.so")


_lib = None


def lib():
    """Load libblubcore.so; fails loudly when it has not been built (``python -m blub_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise BlubError(f"{path} is missing: build it with `python -m blub_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u32, f3 = C.c_void_p, C.c_uint32, C.POINTER(C.c_float)
    sig = {
        "blub_fluid_create": (C.c_int, [C.POINTER(vp), u32, u32, u32, u32, C.c_int, vp]),
        "blub_fluid_destroy": (None, [vp]),
        "blub_fluid_create_slab": (C.c_int, [C.POINTER(vp), u32, u32, u32, u32, C.c_int, vp, C.c_int, C.c_int]),
        "blub_fluid_slab_window": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "blub_fluid_attach_slab_peers": (C.c_int, [vp, C.POINTER(vp), C.c_int]),
        "blub_fluid_slab_error": (C.c_int, [vp]),
        "blub_solid_voxelize": (C.c_int, [vp, C.POINTER(u32), C.POINTER(RigidObject), C.c_float, f3, C.c_double, C.c_double, C.c_int, vp,
                                          C.POINTER(RigidState)]),
        "blub_simulation_delta_ns": (C.c_uint64, [C.c_uint64]),
        "blub_duration_as_secs_f32": (C.c_float, [C.c_uint64]),
        "blub_timer_steps_in_frame": (u32, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64]),
        "blub_fluid_stream": (vp, [vp]),
        "blub_device_malloc": (C.c_int, [C.POINTER(vp), C.c_size_t, C.c_int]),
        "blub_device_free": (C.c_int, [vp]),
        "blub_mesh_create": (C.c_int, [C.POINTER(vp), vp, u32, vp, u32, C.c_int]),
        "blub_mesh_load_obj": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int]),
        "blub_mesh_destroy": (None, [vp]),
        "blub_mesh_info": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32)]),
        "blub_obj_read": (C.c_int, [C.c_char_p, vp, u32, vp, u32, C.POINTER(u32)]),
        "blub_solid_voxelize_mesh": (C.c_int, [vp, C.POINTER(u32), vp, C.POINTER(RigidObject), C.c_float, f3, C.c_double, C.c_double, C.c_int, vp,
                                               C.POINTER(RigidState)]),
        "blub_scene_static_object": (C.c_int, [C.c_char_p, u32, C.POINTER(RigidObject), C.c_char_p, C.c_size_t]),
        "blub_ipc_export": (C.c_int, [vp, C.c_char_p]),
        "blub_ipc_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(vp)]),
        "blub_ipc_close": (C.c_int, [vp]),
        "blub_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
        "blub_fluid_add_cube": (C.c_int, [vp, f3, f3]),
        "blub_fluid_set_gravity_grid": (C.c_int, [vp, f3]),
        "blub_fluid_num_particles": (u32, [vp]),
        "blub_fluid_grid_dimension": (None, [vp, C.POINTER(u32)]),
        "blub_fluid_solver_config": (C.POINTER(SolverConfig), [vp, C.c_int]),
        "blub_fluid_rebinning_frequency": (C.POINTER(u32), [vp]),
        "blub_fluid_solver_stats": (C.c_size_t, [vp, C.c_int, C.POINTER(SolverSample), C.c_size_t]),
        "blub_fluid_update_statistics": (None, [vp]),
        "blub_fluid_set_solid_voxels": (C.c_int, [vp, vp]),
        "blub_fluid_step": (C.c_int, [vp, C.c_double]),
        "blub_fluid_view": (C.c_int, [vp, C.POINTER(FluidView)]),
        "blub_fluid_set_quirks": (C.c_int, [vp, C.POINTER(Quirks)]),
        "blub_fluid_synchronize": (C.c_int, [vp]),
        "blub_last_error": (C.c_char_p, []),
        "blub_version": (C.c_char_p, []),
        "blub_scene_load": (C.c_int, [C.POINTER(vp), C.c_char_p, C.c_int, vp]),
        "blub_scene_info": (C.c_int, [C.c_char_p, C.POINTER(SceneInfo)]),
        "blub_fluid_download": (C.c_int, [vp, C.c_int, vp, C.c_size_t]),
        "blub_fluid_upload": (C.c_int, [vp, C.c_int, vp, C.c_size_t]),
        "blub_fluid_set_particles": (C.c_int, [vp, u32, vp, vp, vp, vp]),
        "blub_fluid_step_stages": (C.c_int, [vp, C.c_double, C.c_int, C.c_int]),
        "blub_fluid_solve_only": (C.c_int, [vp, C.c_int, C.c_double]),
        "blub_fluid_last_solve": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
        "blub_fluid_solver_work": (C.c_int, [vp, C.POINTER(u32)]),
        "blub_fluid_time_solve": (C.c_int, [vp, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_float)]),
        "blub_fluid_time_steps": (C.c_int, [vp, C.c_double, C.c_int, C.POINTER(C.c_float)]),
        "blub_fluid_step_timed": (C.c_int, [vp, C.c_double, C.POINTER(C.c_float)]),
        "blub_fluid_set_graph_replay": (C.c_int, [vp, C.c_int]),
        "blub_fluid_set_solver_path": (C.c_int, [vp, C.c_int]),
        "blub_fluid_set_transfer_path": (C.c_int, [vp, C.c_int]),
        "blub_kernel_launch_count": (C.c_uint64, [C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc, allow=(BLUB_OK,)):
    if rc not in allow:
        raise BlubError(f"libblubcore error {rc}: {lib().blub_last_error().decode()}")
    return rc


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def ipc_open(handle: bytes, device: int) -> int:
    out = C.c_void_p()
    _check(lib().blub_ipc_open(handle, device, C.byref(out)))
    return int(out.value)


def enable_peer_access(device: int, peer: int):
    _check(lib().blub_enable_peer_access(device, peer))


def solid_voxelize(rgba16f_device_ptr, dims, obj, scale, fluid_world_position, t, dt, clear_first=True, cuda_stream=None):
    """Write one analytic rigid solid (dict, see RigidObject.from_dict) into an RGBA16F device volume; returns its RigidState."""
    d = (C.c_uint32 * 3)(*[int(v) for v in dims])
    st = RigidState()
    _check(lib().blub_solid_voxelize(rgba16f_device_ptr, d, C.byref(RigidObject.from_dict(obj)), float(scale), _f3(fluid_world_position), float(t),
                                     float(dt), 1 if clear_first else 0, cuda_stream, C.byref(st)))
    return st


def obj_read(path):
    """Positions [nv, 3] float32 and triangles [nt, 3] uint32 of a Wavefront OBJ, read by the library's host-side loader."""
    import numpy as np
    counts = (C.c_uint32 * 2)()
    _check(lib().blub_obj_read(os.fsencode(path), None, 0, None, 0, counts))
    pos = np.zeros((counts[0], 3), dtype=np.float32)
    idx = np.zeros(counts[1], dtype=np.uint32)
    _check(lib().blub_obj_read(os.fsencode(path), pos.ctypes.data_as(C.c_void_p), counts[0], idx.ctypes.data_as(C.c_void_p), counts[1], counts))
    return pos, idx.reshape(-1, 3)


class Mesh:
    """Device-resident triangle mesh for the hull voxelizer (MeshVertices / MeshIndices of the reference's voxelization pass)."""

    def __init__(self, positions=None, triangles=None, obj_path=None, device=0):
        import numpy as np
        self.h = C.c_void_p()
        if obj_path is not None:
            _check(lib().blub_mesh_load_obj(C.byref(self.h), os.fsencode(obj_path), device))
        else:
            pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
            idx = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1)
            _check(lib().blub_mesh_create(C.byref(self.h), pos.ctypes.data_as(C.c_void_p), pos.shape[0], idx.ctypes.data_as(C.c_void_p), idx.shape[0], device))

    def info(self):
        nv, nt = C.c_uint32(), C.c_uint32()
        _check(lib().blub_mesh_info(self.h, C.byref(nv), C.byref(nt)))
        return int(nv.value), int(nt.value)

    def voxelize(self, rgba16f_device_ptr, dims, placement, scale, fluid_world_position, t, dt, clear_first=True, cuda_stream=None):
        """One draw of the reference's voxelization pass for this mesh (placement: dict as for RigidObject.from_dict, shape ignored)."""
        d = (C.c_uint32 * 3)(*[int(v) for v in dims])
        st = RigidState()
        obj = placement if isinstance(placement, RigidObject) else RigidObject.from_dict(dict(placement, shape="mesh"))
        _check(lib().blub_solid_voxelize_mesh(rgba16f_device_ptr, d, self.h, C.byref(obj), float(scale), _f3(fluid_world_position), float(t), float(dt),
                                              1 if clear_first else 0, cuda_stream, C.byref(st)))
        return st

    def close(self):
        if self.h:
            lib().blub_mesh_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def scene_static_object(path, index):
    """(model path, RigidObject placement) of static_objects[index] of a scene file."""
    obj = RigidObject()
    buf = C.create_string_buffer(1024)
    _check(lib().blub_scene_static_object(os.fsencode(path), index, C.byref(obj), buf, len(buf)))
    return buf.value.decode(), obj


def kernel_launch_count(reset=False) -> int:
    return int(lib().blub_kernel_launch_count(1 if reset else 0))


def scene_info(path):
    info = SceneInfo()
    _check(lib().blub_scene_info(os.fsencode(path), C.byref(info)))
    return info


class HybridFluid:
    """HybridFluid::new (hybrid_fluid.rs:92-100)."""

    PARTICLES_PER_GRID_CELL = 8  # hybrid_fluid.rs:90

    def __init__(self, nx, ny, nz, max_num_particles, device=0, cuda_stream=None, _handle=None):
        self.L = lib()
        if _handle is None:
            h = C.c_void_p()
            _check(self.L.blub_fluid_create(C.byref(h), nx, ny, nz, max_num_particles, device, cuda_stream))
            self.h = h
        else:
            self.h = _handle
        d = (C.c_uint32 * 3)()
        self.L.blub_fluid_grid_dimension(self.h, d)
        self.nx, self.ny, self.nz = int(d[0]), int(d[1]), int(d[2])
        self.n = self.nx * self.ny * self.nz

    @classmethod
    def create_slab(cls, nx, ny, nz_owned, max_num_particles, rank, world, device=0, cuda_stream=None):
        """Rank `rank` of a z-slab decomposition (4 ghost planes on both sides of the nz_owned owned planes)."""
        L = lib()
        h = C.c_void_p()
        _check(L.blub_fluid_create_slab(C.byref(h), nx, ny, nz_owned, max_num_particles, device, cuda_stream, rank, world))
        return cls(0, 0, 0, 0, _handle=h)

    def slab_window(self):
        w, n = C.c_void_p(), C.c_size_t()
        _check(self.L.blub_fluid_slab_window(self.h, C.byref(w), C.byref(n)))
        return int(w.value), int(n.value)

    def attach_slab_peers(self, windows):
        arr = (C.c_void_p * len(windows))(*windows)
        _check(self.L.blub_fluid_attach_slab_peers(self.h, arr, len(windows)))

    def slab_error(self):
        return int(self.L.blub_fluid_slab_error(self.h))

    def ipc_export_window(self):
        buf = C.create_string_buffer(64)
        _check(self.L.blub_ipc_export(self.slab_window()[0], buf))
        return bytes(buf.raw)

    @classmethod
    def from_scene(cls, path, device=0, cuda_stream=None):
        """Scene::new + create_fluid_from_config (src/scene/mod.rs:56-144) on an unchanged scene JSON."""
        L = lib()
        h = C.c_void_p()
        _check(L.blub_scene_load(C.byref(h), os.fsencode(path), device, cuda_stream))
        return cls(0, 0, 0, 0, _handle=h)

    def close(self):
        if getattr(self, "h", None):
            self.L.blub_fluid_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference surface --------------------------------------------------------------------
    def add_fluid_cube(self, min_grid, max_grid):
        rc = _check(self.L.blub_fluid_add_cube(self.h, _f3(min_grid), _f3(max_grid)), allow=(BLUB_OK, BLUB_WARN_TRUNCATED))
        return rc == BLUB_WARN_TRUNCATED

    def set_gravity_grid(self, g):
        _check(self.L.blub_fluid_set_gravity_grid(self.h, _f3(g)))

    @property
    def num_particles(self):
        return int(self.L.blub_fluid_num_particles(self.h))

    def grid_dimension(self):
        return self.nx, self.ny, self.nz

    def pressure_solver_config_velocity(self):
        return self.L.blub_fluid_solver_config(self.h, 0).contents

    def pressure_solver_config_density(self):
        return self.L.blub_fluid_solver_config(self.h, 1).contents

    def set_solver_config(self, which, error_tolerance=0.1, max_num_iterations=32, error_check_frequency=4):
        c = self.L.blub_fluid_solver_config(self.h, which).contents
        c.error_tolerance, c.max_num_iterations, c.error_check_frequency = error_tolerance, max_num_iterations, error_check_frequency

    def set_rebin_frequency(self, f):
        self.L.blub_fluid_rebinning_frequency(self.h)[0] = int(f)

    def pressure_solver_stats(self, which):
        buf = (SolverSample * 100)()
        n = self.L.blub_fluid_solver_stats(self.h, which, buf, 100)
        return [(buf[k].error, buf[k].iteration_count) for k in range(n)]

    def update_statistics(self):
        self.L.blub_fluid_update_statistics(self.h)

    def set_solid_voxels(self, device_ptr):
        _check(self.L.blub_fluid_set_solid_voxels(self.h, device_ptr))

    def step(self, dt=DT_120HZ):
        _check(self.L.blub_fluid_step(self.h, dt))

    def view(self):
        v = FluidView()
        _check(self.L.blub_fluid_view(self.h, C.byref(v)))
        return v

    def set_quirks(self, precond_mode=0):
        q = Quirks()
        q.precond_mode = precond_mode
        _check(self.L.blub_fluid_set_quirks(self.h, C.byref(q)))

    def synchronize(self):
        _check(self.L.blub_fluid_synchronize(self.h))

    # -- taps ---------------------------------------------------------------------------------
    def download_grid(self, tap):
        dt = np.int8 if tap == TAP_MARKER else np.float32
        out = np.empty((self.nz, self.ny, self.nx), dtype=dt)
        _check(self.L.blub_fluid_download(self.h, tap, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload_grid(self, tap, arr):
        dt = np.int8 if tap == TAP_MARKER else np.float32
        arr = np.ascontiguousarray(arr, dtype=dt).reshape(self.nz, self.ny, self.nx)
        _check(self.L.blub_fluid_upload(self.h, tap, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def download_particles(self, tap=TAP_POS):
        out = np.empty((self.num_particles, 4), dtype=np.float32)
        if out.size:
            _check(self.L.blub_fluid_download(self.h, tap, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def set_particles(self, pos4, rowx=None, rowy=None, rowz=None):
        pos4 = np.ascontiguousarray(pos4, dtype=np.float32).reshape(-1, 4)
        rows = [None if r is None else np.ascontiguousarray(r, dtype=np.float32).reshape(-1, 4) for r in (rowx, rowy, rowz)]
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        _check(self.L.blub_fluid_set_particles(self.h, pos4.shape[0], ptr(pos4), ptr(rows[0]), ptr(rows[1]), ptr(rows[2])))

    def step_stages(self, dt, frm, to):
        _check(self.L.blub_fluid_step_stages(self.h, dt, frm, to))

    def solve_only(self, which, dt=DT_120HZ):
        _check(self.L.blub_fluid_solve_only(self.h, which, dt))

    def last_solve(self, which):
        e, it = C.c_float(0), C.c_int32(0)
        _check(self.L.blub_fluid_last_solve(self.h, which, C.byref(e), C.byref(it)))
        return float(e.value), int(it.value)

    def solver_work(self):
        """{tiles, columns, cells_per_tile, cells_per_column} of the most recent solve's work lists."""
        out = (C.c_uint32 * 4)()
        _check(self.L.blub_fluid_solver_work(self.h, out))
        return {"tiles": int(out[0]), "columns": int(out[1]), "cells_per_tile": int(out[2]), "cells_per_column": int(out[3])}

    def time_solve(self, which, dt, repetitions):
        ms = (C.c_float * repetitions)()
        _check(self.L.blub_fluid_time_solve(self.h, which, dt, repetitions, ms))
        return [float(x) for x in ms]

    def step_timed(self, dt=DT_120HZ):
        ms = (C.c_float * 14)()
        _check(self.L.blub_fluid_step_timed(self.h, dt, ms))
        return [float(x) for x in ms]

    def set_solver_path(self, persistent):
        """True / 1: persistent cooperative PCG, dense tiles + column list (default); False / 0: three kernels per iteration;
        2 / "tma": TMA-staged tiles; 4 / "dense": tile kernel without the sparsity skip; 6 / "tiles": tile kernel with the sparse tile bodies."""
        names = {"tma": 2, "dense": 4, "tiles": 6}
        mode = names[persistent] if persistent in names else (int(persistent) if persistent not in (True, False) else (1 if persistent else 0))
        _check(self.L.blub_fluid_set_solver_path(self.h, mode))

    def set_transfer_path(self, scatter):
        """True / "scatter" (default): warp-aggregated atomic scatter; False / "gather": deterministic gather P2G over per-step cell lists."""
        _check(self.L.blub_fluid_set_transfer_path(self.h, 1 if scatter in (True, 1, "scatter") else 0))

    def stream(self):
        """The CUDA stream (as an integer handle) the fluid's work is enqueued on."""
        return self.L.blub_fluid_stream(self.h)

    def set_graph_replay(self, enabled):
        _check(self.L.blub_fluid_set_graph_replay(self.h, 1 if enabled else 0))

    def time_steps(self, dt, steps):
        ms = C.c_float(0)
        _check(self.L.blub_fluid_time_steps(self.h, dt, steps, C.byref(ms)))
        return float(ms.value)
