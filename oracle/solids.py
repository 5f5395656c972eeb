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


# ------------------------------------------------------------------------------------------------ triangle meshes (hull voxelization)
def mesh_pose(obj, scale, fluid_world_position, t, dt):
    """MeshDataGpu (src/scene/models.rs:186-224) in float32: 3x4 VoxelTransform rows, centre, scaled axis, velocity."""
    st = rigid_state(obj, scale, fluid_world_position, t, dt)
    k = f32(f32(obj.get("scale", 1.0)) / f32(scale))
    m = np.zeros((3, 4), dtype=f32)
    m[:, :3] = st["R"].astype(f32) * k
    m[:, 3] = st["centre"].astype(f32)
    return {"m": m, "centre": st["centre"].astype(f32), "axis": st["axis"].astype(f32), "velocity": st["velocity"].astype(f32)}


def _swz(side, v):
    return v[[2, 1, 0]] if side == 0 else (v[[0, 2, 1]] if side == 1 else v)


def _speed(pose, pos):
    """ComputeVoxelSpeed (conservative_hull.frag:17-23) on an [n, 3] float32 array, same operation order as voxelize_core.hpp."""
    a, c, v = pose["axis"], pose["centre"], pose["velocity"]
    p = pos - c
    pa = (p[:, 0] * a[0] + p[:, 1] * a[1]) + p[:, 2] * a[2]
    q = p - pa[:, None] * a
    return np.stack([(a[1] * q[:, 2] - a[2] * q[:, 1]) + v[0], (a[2] * q[:, 0] - a[0] * q[:, 2]) + v[1], (a[0] * q[:, 1] - a[1] * q[:, 0]) + v[2]], axis=1)


def voxelize_mesh_hull(positions, indices, pose, dims):
    """The reference's voxelization pass (src/scene/voxelization.rs:118-157, shader/voxelize/conservative_hull.{vert,frag}) for one mesh,
    float32 throughout, with the driver-defined parts fixed as in blub_b200/csrc/voxelize_core.hpp (closed-square conservative
    coverage, plane depth at the pixel centre clamped to the triangle's range, sequential store order).
    Returns ([nz, ny, nx, 4] float32 volume, [nz, ny, nx] int64 owner = 4 * triangle + store kind + 1, 0 where nothing was written)."""
    nx, ny, nz = dims
    res_hi = np.array([nx - 1, ny - 1, nz - 1], dtype=f32)
    size = int(max(dims))
    lim = f32(size)
    out = np.zeros((nz, ny, nx, 4), dtype=f32)
    owner = np.zeros((nz, ny, nx), dtype=np.int64)
    pos = np.asarray(positions, dtype=f32).reshape(-1, 3)
    m = pose["m"]
    vv = ((m[:, 0] * pos[:, 0:1] + m[:, 1] * pos[:, 1:2]) + m[:, 2] * pos[:, 2:3]) + m[:, 3]  # [nv, 3]
    tris = np.asarray(indices, dtype=np.int64).reshape(-1, 3)
    for ti, (ia, ib, ic) in enumerate(tris):
        a, b, c = vv[ia], vv[ib], vv[ic]
        e1, e2 = b - a, c - a
        n = np.abs(np.array([e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]], dtype=f32))
        side = 0 if n[0] > n[1] else 1
        side = side if n[side] > n[2] else 2
        P = np.stack([_swz(side, a), _swz(side, b), _swz(side, c)])
        u, w = P[1] - P[0], P[2] - P[0]
        nzc = u[0] * w[1] - u[1] * w[0]
        if not abs(nzc) > 0:
            continue
        gx = -(u[1] * w[2] - u[2] * w[1]) / nzc
        gy = -(u[2] * w[0] - u[0] * w[2]) / nzc
        zmin, zmax = P[:, 2].min(), P[:, 2].max()
        xmin, xmax, ymin, ymax = P[:, 0].min(), P[:, 0].max(), P[:, 1].min(), P[:, 1].max()
        if xmax < 0 or ymax < 0 or xmin > lim or ymin > lim or zmax < 0 or zmin > lim:
            continue
        x0, y0 = int(np.floor(max(xmin, f32(0)))), int(np.floor(max(ymin, f32(0))))
        x1, y1 = int(np.floor(min(xmax, lim - f32(1)))), int(np.floor(min(ymax, lim - f32(1))))
        if x0 > 0 and f32(x0) == xmin:
            x0 -= 1
        if y0 > 0 and f32(y0) == ymin:
            y0 -= 1
        if x1 < x0 or y1 < y0:
            continue
        py, px = np.meshgrid(np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
        px, py = px.ravel(), py.ravel()
        X, Y = px.astype(f32), py.astype(f32)
        keep = np.ones(px.shape, dtype=bool)
        orient = f32(1.0) if nzc > 0 else f32(-1.0)
        for k in range(3):
            q0, q1, q2 = P[k], P[(k + 1) % 3], P[(k + 2) % 3]
            ex, ey = -(q1[1] - q0[1]) * orient, (q1[0] - q0[0]) * orient
            eo = -(ex * q0[0] + ey * q0[1])
            eh = (ex * q2[0] + ey * q2[1]) + eo
            hi = (ex * (X + f32(1) if ex >= 0 else X) + ey * (Y + f32(1) if ey >= 0 else Y)) + eo
            lo = (ex * (X if ex >= 0 else X + f32(1)) + ey * (Y if ey >= 0 else Y + f32(1))) + eo
            keep &= ~((hi < 0) | (lo > eh))
        px, py = px[keep], py[keep]
        fx, fy = px.astype(f32) + f32(0.5), py.astype(f32) + f32(0.5)
        z = (P[0][2] + gx * (fx - P[0][0])) + gy * (fy - P[0][1])
        z = np.minimum(np.maximum(z, zmin), zmax)
        ok = ~((z < 0) | (z > lim))
        fx, fy, z = fx[ok], fy[ok], z[ok]
        mc = max(abs(gx), abs(gy))
        stores = [(0, np.stack([np.trunc(fx), np.trunc(fy), np.trunc(z)], 1), np.ones(z.shape, bool)),
                  (1, np.stack([fx, fy, z - f32(1)], 1), np.floor(z) != np.floor(z - mc)),
                  (2, np.stack([fx, fy, z + f32(1)], 1), np.floor(z) != np.floor(z + mc))]
        # sequential semantics: fragment by fragment, store by store.  Within one triangle the stores of different fragments can hit
        # the same voxel; priority = (triangle, kind), and among equal priorities every candidate carries the same cell, so only the
        # velocity could differ -- resolved towards the LAST fragment in raster order (x fastest), as the CUDA resolve pass does.
        for kind, sw, take in stores:
            sw = sw[take]
            if sw.shape[0] == 0:
                continue
            p3 = sw[:, [2, 1, 0]] if side == 0 else (sw[:, [0, 2, 1]] if side == 1 else sw)
            p3 = np.minimum(np.maximum(p3, f32(0)), res_hi)
            cell = p3.astype(np.int64)
            vel = _speed(pose, p3)
            prio = 4 * ti + kind + 1
            for (cx, cy, cz), v3 in zip(cell, vel):
                if owner[cz, cy, cx] <= prio:
                    owner[cz, cy, cx] = prio
                    out[cz, cy, cx, :3] = v3
                    out[cz, cy, cx, 3] = 1.0
    return out, owner


def read_obj(path):
    """Positions + fan-triangulated faces of a Wavefront OBJ (what tobj::load_obj(triangulate, ignore_points, ignore_lines) keeps of the
    geometry, src/scene/models.rs:252-262); 1-based and negative (relative) indices, v / v/vt / v//vn / v/vt/vn corners."""
    verts, tris = [], []
    with open(path) as fh:
        for line in fh:
            line = line.split("#", 1)[0].split()
            if not line:
                continue
            if line[0] == "v":
                verts.append([float(x) for x in line[1:4]])
            elif line[0] == "f":
                idx = []
                for corner in line[1:]:
                    i = int(corner.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, dtype=f32).reshape(-1, 3), np.asarray(tris, dtype=np.uint32).reshape(-1, 3)
