"""Triangle-mesh hull voxelizer: the reference's conservative-rasterization pass
(src/scene/voxelization.rs:118-157, shader/voxelize/conservative_hull.{vert,frag}) as CUDA kernels.

CPU (no GPU needed): the per-triangle / per-pixel functions the kernels run live in blub_b200/csrc/voxelize_core.hpp; a host twin built
from that very header (tests/native/voxcore_host.cpp) must agree BIT FOR BIT with the NumPy restatement (oracle/solids.py), and both
must satisfy float64 geometric bounds that do not share any code with them.  GPU: the CUDA voxelizer against the same restatement.
(The file sorts last on purpose: these kernels are the newest code of the repository.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from blub_b200 import fluid as F
from oracle import solids as S
from tests import util
from tests.util import DT

NATIVE = os.path.join(util.HERE, "native")


# ------------------------------------------------------------------------------------------------ meshes
def box_mesh(half):
    v = np.array([[x, y, z] for z in (-1, 1) for y in (-1, 1) for x in (-1, 1)], dtype=np.float32) * np.asarray(half, dtype=np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = []
    for q in quads:
        tris += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
    return v, np.array(tris, dtype=np.uint32)


def icosphere(radius, subdivisions=2):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.array(v) * radius).astype(np.float32), np.array(f, dtype=np.uint32)


PLACEMENT = {"world_position": [0.16, 0.17, 0.15], "scale": 1.0, "rotation_angles": [20.0, 35.0, 10.0],
             "rotation": {"axis": [0.0, 1.0, 0.3], "deg_per_sec": 40.0},
             "translation": {"target": [0.3, 0.17, 0.15], "curve": "SmoothStep", "duration": 2.0}}
SCALE, ORIGIN = 0.01, (0.0, 0.0, 0.0)
CASES = [("box", (32, 32, 32), 0.0), ("box", (32, 32, 32), 0.7), ("box", (64, 32, 32), 2.9), ("sphere", (32, 32, 32), 1.3), ("sphere", (64, 32, 32), 0.4)]


def mesh_of(kind):
    return box_mesh([0.05, 0.08, 0.04]) if kind == "box" else icosphere(0.07)


# ------------------------------------------------------------------------------------------------ host twin
@pytest.fixture(scope="module")
def voxcore():
    so = os.path.join(NATIVE, "libvoxcore_host.so")
    src = [os.path.join(NATIVE, "voxcore_host.cpp"), os.path.join(util.HERE, "..", "blub_b200", "csrc", "voxelize_core.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src[0]])
    return C.CDLL(so)


def host_voxelize(lib, positions, tris, pose, dims):
    nx, ny, nz = dims
    rgba = np.zeros((nz, ny, nx, 4), np.float32)
    owner = np.zeros((nz, ny, nx), np.int64)
    p21 = np.concatenate([pose["m"].ravel(), pose["centre"], pose["axis"], pose["velocity"]]).astype(np.float32)
    v = np.ascontiguousarray(positions, np.float32)
    t = np.ascontiguousarray(tris, np.uint32)
    rc = lib.voxcore_host_voxelize(v.ctypes.data_as(C.c_void_p), C.c_uint32(len(v)), t.ctypes.data_as(C.c_void_p), C.c_uint32(len(t)),
                                   p21.ctypes.data_as(C.c_void_p), (C.c_int32 * 3)(*dims), rgba.ctypes.data_as(C.c_void_p), owner.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return rgba, owner


@pytest.mark.parametrize("kind,dims,t", CASES)
def test_kernel_arithmetic_matches_numpy_restatement_bit_for_bit(voxcore, kind, dims, t):
    v, tris = mesh_of(kind)
    pose = S.mesh_pose(PLACEMENT, SCALE, ORIGIN, t, DT)
    want, want_owner = S.voxelize_mesh_hull(v, tris, pose, dims)
    got, got_owner = host_voxelize(voxcore, v, tris, pose, dims)
    assert want[..., 3].sum() > 300
    assert np.array_equal(got_owner >> 32, want_owner)
    assert np.array_equal(got, want)


def point_triangle_distance(p, a, b, c):
    """Distance from points p [n, 3] to triangle abc, float64 (Ericson, Real-Time Collision Detection 5.1.5: closest point)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ap @ ab, ap @ ac
    bp = p - b
    d3, d4 = bp @ ab, bp @ ac
    cp = p - c
    d5, d6 = cp @ ab, cp @ ac
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    out = np.empty_like(p)
    denom = np.where(va + vb + vc == 0, 1.0, va + vb + vc)
    vv, ww = vb / denom, vc / denom
    out[:] = a + ab * vv[:, None] + ac * ww[:, None]  # interior
    with np.errstate(divide="ignore", invalid="ignore"):
        m = (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0)
        w_bc = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        out[m] = (b + (c - b) * w_bc[:, None])[m]
        m = (vb <= 0) & (d2 >= 0) & (d6 <= 0)
        out[m] = (a + ac * (d2 / (d2 - d6))[:, None])[m]
        m = (vc <= 0) & (d1 >= 0) & (d3 <= 0)
        out[m] = (a + ab * (d1 / (d1 - d3))[:, None])[m]
    m = (d6 >= 0) & (d5 <= d6)
    out[m] = c
    m = (d3 >= 0) & (d4 <= d3)
    out[m] = b
    m = (d1 <= 0) & (d2 <= 0)
    out[m] = a
    return np.linalg.norm(p - out, axis=1)


@pytest.mark.parametrize("kind,dims,t", [CASES[1], CASES[4]])  # poses that keep the whole mesh inside the grid
def test_hull_satisfies_geometric_bounds(kind, dims, t):
    """Independent float64 sandwich: (1) every voxel that contains a point of the surface is marked (conservative coverage);
    (2) every marked voxel lies within 2.5 cells of the surface (main voxel: < sqrt(3)/2 + 1 for the depth rounding; +-1 layer: one more)."""
    v, tris = mesh_of(kind)
    pose = S.mesh_pose(PLACEMENT, SCALE, ORIGIN, t, DT)
    vol, _ = S.voxelize_mesh_hull(v, tris, pose, dims)
    marked = vol[..., 3] > 0
    m = pose["m"].astype(np.float64)
    vv = v.astype(np.float64) @ m[:, :3].T + m[:, 3]
    rng = np.random.default_rng(1)
    for a, b, c in vv[tris.astype(np.int64)]:
        r1, r2 = rng.random(300), rng.random(300)
        s = np.sqrt(r1)
        pts = (1 - s)[:, None] * a + (s * (1 - r2))[:, None] * b + (s * r2)[:, None] * c
        cells = np.floor(pts).astype(int)
        assert marked[cells[:, 2], cells[:, 1], cells[:, 0]].all()
    zz, yy, xx = np.nonzero(marked)
    centres = np.stack([xx, yy, zz], axis=1) + 0.5
    dist = np.full(len(centres), np.inf)
    for a, b, c in vv[tris.astype(np.int64)]:
        dist = np.minimum(dist, point_triangle_distance(centres, a, b, c))
    assert dist.max() <= 2.5, dist.max()


def test_velocity_field_of_a_rotating_translating_mesh():
    """ComputeVoxelSpeed (conservative_hull.frag:17-23): for a UNIT rotation speed (|a| = 1 rad/s) the formula is the rigid-body
    velocity v + a x (p - centre) evaluated at the stored voxel position."""
    v, tris = box_mesh([0.05, 0.08, 0.04])
    placement = dict(PLACEMENT, rotation={"axis": [0.0, 2.0, 0.0], "deg_per_sec": float(np.rad2deg(1.0))})
    pose = S.mesh_pose(placement, SCALE, ORIGIN, 0.9, DT)
    assert abs(np.linalg.norm(pose["axis"]) - 1.0) < 1e-6
    vol, owner = S.voxelize_mesh_hull(v, tris, pose, (32, 32, 32))
    zz, yy, xx = np.nonzero((owner > 0) & ((owner - 1) % 4 == 0))  # main stores: velocity at the integer voxel position
    p = np.stack([xx, yy, zz], axis=1).astype(np.float64) - pose["centre"]
    want = np.cross(pose["axis"].astype(np.float64), p) + pose["velocity"]
    assert np.abs(vol[zz, yy, xx, :3] - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


# ------------------------------------------------------------------------------------------------ OBJ / scene surface (host)
OBJ_TEXT = """# quad, triangle with relative indices, a line and a point (ignored)
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vn 0 0 1
vt 0.5 0.5
f 1/1/1 2/1/1 3/1/1 4/1/1
v 0 0 1
f -1 -2// -3//1
l 1 2
p 1
"""


def test_obj_reader(tmp_path):
    path = tmp_path / "mesh.obj"
    path.write_text(OBJ_TEXT)
    pos, tris = F.obj_read(str(path))
    assert pos.tolist() == [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1]]
    assert tris.tolist() == [[0, 1, 2], [0, 2, 3], [4, 3, 2]]  # fan triangulation; -1 = the vertex defined last
    ref_pos, ref_tris = S.read_obj(str(path))
    assert np.array_equal(pos, ref_pos) and np.array_equal(tris, ref_tris)
    for bad in ("v 0 0\n", "v 0 0 0\nf 1 2 3\n", "v 0 0 0\nf 0 1 1\n", "v 0 0 0\nf a b c\n"):
        path.write_text(bad)
        with pytest.raises(F.BlubError):
            F.obj_read(str(path))
    with pytest.raises(F.BlubError):
        F.obj_read(str(tmp_path / "missing.obj"))


def test_scene_static_objects_are_parsed(tmp_path):
    scene = {"gravity": {"x": 0, "y": -9.81, "z": 0},
             "fluid": {"world_position": {"x": 0, "y": 0, "z": 0}, "grid_to_world_scale": 0.01, "grid_dimension": {"x": 32, "y": 32, "z": 32},
                       "max_num_particles": 1000, "fluid_cubes": []},
             "static_objects": [
                 {"model": "wgpu-logo/wgpu.obj", "world_position": {"x": 0.0, "y": 0.0, "z": 0.32}, "scale": 0.12,
                  "rotation_angles": {"x": 0.0, "y": 30.0, "z": 0.0},
                  "animation": {"rotation": {"axis": {"x": 0.0, "y": 5.0, "z": 0.0}, "deg_per_sec": 180.0},
                                "translation": {"target": {"x": 1.28, "y": 0.0, "z": 0.32}, "curve": "SmoothStep", "duration": 2.0}}},
                 {"model": "cube.obj", "world_position": {"x": 1, "y": 2, "z": 3}, "scale": 2.0, "rotation_angles": {"x": 1, "y": 2, "z": 3}}]}
    import json
    path = tmp_path / "scene.json"
    path.write_text(json.dumps(scene))
    assert F.scene_info(str(path)).num_static_objects == 2
    model, o = F.scene_static_object(str(path), 0)
    assert model == "wgpu-logo/wgpu.obj" and o.shape == 2 and abs(o.scale - 0.12) < 1e-7
    assert list(o.rotation_angles_deg) == [0.0, 30.0, 0.0] and o.has_rotation == 1 and o.has_translation == 1
    assert list(o.rotation_axis) == [0.0, 5.0, 0.0] and o.rotation_deg_per_sec == 180.0
    assert o.translation_curve == 1 and o.translation_duration == 2.0 and abs(o.translation_target[0] - 1.28) < 1e-6
    model, o = F.scene_static_object(str(path), 1)
    assert model == "cube.obj" and o.has_rotation == 0 and o.has_translation == 0 and list(o.world_position) == [1.0, 2.0, 3.0]
    with pytest.raises(F.BlubError):
        F.scene_static_object(str(path), 2)
    scene["static_objects"][0]["animation"]["translation"]["curve"] = "Cubic"
    path.write_text(json.dumps(scene))
    with pytest.raises(F.BlubError):
        F.scene_info(str(path))


# ------------------------------------------------------------------------------------------------ CUDA
@pytest.mark.gpu
@pytest.mark.parametrize("kind,dims,t", CASES)
def test_cuda_mesh_voxelizer_matches_numpy_restatement(kind, dims, t):
    import torch
    v, tris = mesh_of(kind)
    nx, ny, nz = dims
    mesh = F.Mesh(v, tris)
    assert mesh.info() == (len(v), len(tris))
    vol = torch.full((nz, ny, nx, 4), 7.0, dtype=torch.float16, device="cuda")
    st = mesh.voxelize(vol.data_ptr(), dims, PLACEMENT, SCALE, ORIGIN, t, DT)
    torch.cuda.synchronize()
    got = vol.float().cpu().numpy()
    # the restatement runs on the pose the library evaluated (its float32 trigonometry is compared separately in test_gpu_solids.py)
    k = np.float32(np.float32(PLACEMENT["scale"]) / np.float32(SCALE))
    m = np.zeros((3, 4), dtype=np.float32)
    m[:, :3] = np.array(st.rotation, dtype=np.float32).reshape(3, 3) * k
    m[:, 3] = np.array(st.centre_voxel, dtype=np.float32)
    pose = {"m": m, "centre": np.array(st.centre_voxel, dtype=np.float32), "axis": np.array(st.axis_scaled, dtype=np.float32),
            "velocity": np.array(st.velocity_voxel, dtype=np.float32)}
    ref = S.mesh_pose(PLACEMENT, SCALE, ORIGIN, t, DT)
    assert np.abs(pose["m"] - ref["m"]).max() <= 2e-3 and np.abs(pose["velocity"] - ref["velocity"]).max() <= 2e-2 * max(1.0, np.abs(ref["velocity"]).max())
    want, _ = S.voxelize_mesh_hull(v, tris, pose, dims)
    assert np.array_equal(got[..., 3] > 0, want[..., 3] > 0)
    assert np.array_equal(got, want.astype(np.float16).astype(np.float32))  # same arithmetic, rounded once to fp16
    # a second mesh without clearing overwrites only where it draws
    v2, t2 = box_mesh([0.02, 0.02, 0.02])
    small = F.Mesh(v2, t2)
    small.voxelize(vol.data_ptr(), dims, dict(PLACEMENT, world_position=[0.05, 0.05, 0.05], translation=None, rotation=None), SCALE, ORIGIN, t, DT,
                   clear_first=False)
    torch.cuda.synchronize()
    both = vol.float().cpu().numpy()
    assert (both[..., 3] > 0).sum() > (got[..., 3] > 0).sum() and ((both[..., 3] > 0) >= (got[..., 3] > 0)).all()
    mesh.close()
    small.close()
