"""NumPy restatement of the reference's rigid-object animation and per-voxel solid velocity (test infrastructure only).

    StaticMeshData::{world_position_at_time, rotation_at_time, to_gpu}   src/scene/models.rs:154-224
    ComputeVoxelSpeed                                                    shader/voxelize/conservative_hull.frag:17-23
cgmath 0.18's Euler -> Quaternion conversion is third-party (not vendored): restated from its published source, unverifiable offline.
The reference rasterizes mesh HULLS; the analytic solids used here mark the full interior of a box / sphere instead.
"""
import numpy as np

f32 = np.float32


def world_position_at_time(obj, t):
    pos = np.asarray(obj["world_position"], dtype=f32)
    tr = obj.get("translation")
    if not tr:
        return pos
    dur = f32(tr["duration"])
    prog = f32(np.fmod(f32(t), dur * f32(2.0)))
    if prog > dur:
        prog = dur * f32(2.0) - prog
    prog = f32(prog / dur)
    prog = f32(min(max(prog, f32(0.0)), f32(1.0)))
    if tr.get("curve", "Linear") == "SmoothStep":
        prog = f32(prog * prog * (f32(3.0) - f32(2.0) * prog))
    return pos * (f32(1.0) - prog) + np.asarray(tr["target"], dtype=f32) * prog


def quat_from_euler_deg(deg):
    h = np.deg2rad(np.asarray(deg, dtype=np.float64)) * 0.5
    sx, sy, sz = np.sin(h)
    cx, cy, cz = np.cos(h)
    return np.array([-sx * sy * sz + cx * cy * cz, sx * cy * cz + sy * sz * cx, -sx * sz * cy + sy * cx * cz, sx * sy * cz + sz * cx * cy])


def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def rotation_matrix(obj, t):
    q = quat_from_euler_deg(obj.get("rotation_angles", (0, 0, 0)))
    rot = obj.get("rotation")
    if rot:
        axis = np.asarray(rot["axis"], dtype=np.float64)
        axis = axis / np.linalg.norm(axis)
        ang = np.deg2rad(rot["deg_per_sec"]) * t
        q = quat_mul(q, np.concatenate([[np.cos(ang / 2)], axis * np.sin(ang / 2)]))
    s, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * s), 2 * (x * z + y * s)],
                     [2 * (x * y + z * s), 1 - 2 * (x * x + z * z), 2 * (y * z - x * s)],
                     [2 * (x * z - y * s), 2 * (y * z + x * s), 1 - 2 * (x * x + y * y)]])


def rigid_state(obj, scale, fluid_world_position, t, dt):
    pos = world_position_at_time(obj, t)
    vel = np.zeros(3, dtype=f32)
    if t > dt:  # models.rs:186-191
        vel = (pos - world_position_at_time(obj, f32(t) - f32(dt))) / f32(dt)
    axis_scaled = np.zeros(3)
    rot = obj.get("rotation")
    if rot:
        a = np.asarray(rot["axis"], dtype=np.float64)
        axis_scaled = a / np.linalg.norm(a) * np.deg2rad(rot["deg_per_sec"])
    centre = (pos - np.asarray(fluid_world_position, dtype=f32)) / f32(scale)
    return {"centre": centre.astype(np.float64), "velocity": (vel / f32(scale)).astype(np.float64), "axis": axis_scaled, "R": rotation_matrix(obj, t)}


def voxelize(obj, dims, scale, fluid_world_position, t, dt):
    """[nz, ny, nx, 4] float32 volume: xyz = solid velocity in cells/s, w = 1 inside the box / sphere."""
    nx, ny, nz = dims
    st = rigid_state(obj, scale, fluid_world_position, t, dt)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    p_int = np.stack([x, y, z], axis=-1).astype(np.float64)
    d = p_int + 0.5 - st["centre"]
    half = np.asarray(obj["half_extent"], dtype=np.float64) * obj.get("scale", 1.0) / scale
    if obj.get("shape", "box") == "sphere":
        inside = (d ** 2).sum(-1) <= half[0] ** 2
    else:
        local = d @ st["R"]  # R^T d
        inside = (np.abs(local) <= half).all(-1)
    p = p_int - st["centre"]
    a = st["axis"]
    q = p - (p @ a)[..., None] * a
    v = np.cross(np.broadcast_to(a, q.shape), q) + st["velocity"]
    out = np.zeros((nz, ny, nx, 4), dtype=np.float32)
    out[inside, :3] = v[inside]
    out[inside, 3] = 1.0
    return out, st
