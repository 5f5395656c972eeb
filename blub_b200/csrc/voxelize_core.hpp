// voxelize_core.hpp -- per-triangle / per-pixel arithmetic of the conservative hull voxelizer.
//
// Restates what the reference's voxelization render pass computes (src/scene/voxelization.rs:118-157,
// shader/voxelize/conservative_hull.vert:11-41, conservative_hull.frag:14-52) as plain functions, compiled for the device by
// mesh_voxelizer.cu and -- from this very header -- for the host by tests/native/voxcore_host.cpp, so that a CPU-only test can check
// the arithmetic the kernels run against the NumPy restatement (oracle/solids.py).  No CUDA types in here.
//
// The raster pipeline's driver-defined parts are pinned down as follows (documented, parity unpinned):
//   * conservative rasterization (PrimitiveState::conservative, voxelization.rs:77) = a pixel produces a fragment iff its closed unit
//     square intersects the closed projected triangle (exact separating-axis test; hardware may over-estimate further);
//   * the fragment's depth is the triangle's plane evaluated at the pixel centre, clamped to the triangle's own depth range (pixel
//     centres of conservative fragments can lie outside the triangle), and fragments outside the depth range [0, viewport] are clipped;
//   * dFdxCoarse / dFdyCoarse of the depth are the plane's gradient;
//   * image stores to the same voxel are resolved as if the draw were executed sequentially (triangle order, then the store order
//     inside the fragment shader): one of the orders the unordered stores of the reference allow.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BLUB_HD __host__ __device__ __forceinline__
#else
#define BLUB_HD inline
#endif

namespace blub {
namespace vox {

struct MeshPose {
    float m[12];       // row-major 3x4 VoxelTransform: voxel = m[4r..4r+2] . model + m[4r+3]      (models.rs:196-198,201-203)
    float centre[3];   // VoxelTransform * (0,0,0,1)                                               (conservative_hull.frag:18)
    float axis[3];     // FluidSpaceRotationAxisScaled: unit axis * rad/s                          (models.rs:207-216)
    float velocity[3]; // FluidSpaceVelocity, cells/s                                              (models.rs:206)
};

struct TriSetup {
    int side;           // dominant axis of the normal: 0 = X, 1 = Y, 2 = Z                         (conservative_hull.vert:21-23)
    float p[3][3];      // swizzled vertices (x, y = raster plane, z = depth)                        (:28-36)
    float gx, gy;       // depth plane gradient per pixel
    float zmin, zmax;   // depth range of the triangle
    int x0, x1, y0, y1; // candidate pixel range, clipped to the viewport (inclusive); empty if x1 < x0 or y1 < y0
    float en[3][2];     // edge normals (inside positive)
    float eo[3];        // edge offsets: E_k(q) = en[k] . q + eo[k] >= 0 inside
    float eh[3];        // E_k at the vertex opposite to edge k (the far end of the triangle's projection on that normal)
};

BLUB_HD void transform_vertex(const MeshPose &pose, const float v[3], float out[3]) {
    for (int r = 0; r < 3; ++r) out[r] = ((pose.m[4 * r] * v[0] + pose.m[4 * r + 1] * v[1]) + pose.m[4 * r + 2] * v[2]) + pose.m[4 * r + 3];
}

BLUB_HD void swizzle(int side, const float v[3], float out[3]) { // self-inverse: zyx / xzy / xyz
    if (side == 0) { out[0] = v[2]; out[1] = v[1]; out[2] = v[0]; }
    else if (side == 1) { out[0] = v[0]; out[1] = v[2]; out[2] = v[1]; }
    else { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
}

// returns false for triangles that cannot produce fragments (degenerate or outside the viewport)
BLUB_HD bool setup_triangle(const float a[3], const float b[3], const float c[3], int viewport, TriSetup &t) {
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float n[3] = {fabsf(e1[1] * e2[2] - e1[2] * e2[1]), fabsf(e1[2] * e2[0] - e1[0] * e2[2]), fabsf(e1[0] * e2[1] - e1[1] * e2[0])};
    int side = n[0] > n[1] ? 0 : 1;
    side = n[side] > n[2] ? side : 2;
    t.side = side;
    swizzle(side, a, t.p[0]);
    swizzle(side, b, t.p[1]);
    swizzle(side, c, t.p[2]);
    // depth plane through the swizzled vertices
    const float u[3] = {t.p[1][0] - t.p[0][0], t.p[1][1] - t.p[0][1], t.p[1][2] - t.p[0][2]};
    const float w[3] = {t.p[2][0] - t.p[0][0], t.p[2][1] - t.p[0][1], t.p[2][2] - t.p[0][2]};
    const float nz = u[0] * w[1] - u[1] * w[0]; // twice the signed projected area
    if (!(fabsf(nz) > 0.0f)) return false;
    t.gx = -(u[1] * w[2] - u[2] * w[1]) / nz;
    t.gy = -(u[2] * w[0] - u[0] * w[2]) / nz;
    t.zmin = fminf(t.p[0][2], fminf(t.p[1][2], t.p[2][2]));
    t.zmax = fmaxf(t.p[0][2], fmaxf(t.p[1][2], t.p[2][2]));
    const float xmin = fminf(t.p[0][0], fminf(t.p[1][0], t.p[2][0])), xmax = fmaxf(t.p[0][0], fmaxf(t.p[1][0], t.p[2][0]));
    const float ymin = fminf(t.p[0][1], fminf(t.p[1][1], t.p[2][1])), ymax = fmaxf(t.p[0][1], fmaxf(t.p[1][1], t.p[2][1]));
    const float lim = (float)viewport;
    if (xmax < 0.0f || ymax < 0.0f || xmin > lim || ymin > lim || t.zmax < 0.0f || t.zmin > lim) return false;
    t.x0 = (int)floorf(fmaxf(xmin, 0.0f));
    t.y0 = (int)floorf(fmaxf(ymin, 0.0f));
    t.x1 = (int)floorf(fminf(xmax, lim - 1.0f));
    t.y1 = (int)floorf(fminf(ymax, lim - 1.0f));
    if (t.x0 > 0 && (float)t.x0 == xmin) t.x0 -= 1; // a vertex exactly on a pixel border touches the pixel before it as well
    if (t.y0 > 0 && (float)t.y0 == ymin) t.y0 -= 1;
    const float orient = nz > 0.0f ? 1.0f : -1.0f;
    for (int k = 0; k < 3; ++k) {
        const float *q0 = t.p[k], *q1 = t.p[(k + 1) % 3], *q2 = t.p[(k + 2) % 3];
        t.en[k][0] = -(q1[1] - q0[1]) * orient;
        t.en[k][1] = (q1[0] - q0[0]) * orient;
        t.eo[k] = -(t.en[k][0] * q0[0] + t.en[k][1] * q0[1]);
        t.eh[k] = (t.en[k][0] * q2[0] + t.en[k][1] * q2[1]) + t.eo[k];
    }
    return true;
}

// closed pixel square [px, px+1] x [py, py+1] against the closed triangle: separating axes = the three edge normals (the pixel
// range of setup_triangle already covers the two box axes)
BLUB_HD bool pixel_overlaps(const TriSetup &t, int px, int py) {
    const float x = (float)px, y = (float)py;
    for (int k = 0; k < 3; ++k) {
        const float nx = t.en[k][0], ny = t.en[k][1];
        const float hi = (nx * (nx >= 0.0f ? x + 1.0f : x) + ny * (ny >= 0.0f ? y + 1.0f : y)) + t.eo[k];
        const float lo = (nx * (nx >= 0.0f ? x : x + 1.0f) + ny * (ny >= 0.0f ? y : y + 1.0f)) + t.eo[k];
        if (hi < 0.0f || lo > t.eh[k]) return false;
    }
    return true;
}

BLUB_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// ComputeVoxelSpeed, conservative_hull.frag:17-23 (the axis carries the angular speed; replicated as written)
BLUB_HD void voxel_speed(const MeshPose &pose, const float pos[3], float out[3]) {
    const float p[3] = {pos[0] - pose.centre[0], pos[1] - pose.centre[1], pos[2] - pose.centre[2]};
    const float *a = pose.axis;
    const float pa = (p[0] * a[0] + p[1] * a[1]) + p[2] * a[2];
    const float q[3] = {p[0] - pa * a[0], p[1] - pa * a[1], p[2] - pa * a[2]};
    out[0] = (a[1] * q[2] - a[2] * q[1]) + pose.velocity[0];
    out[1] = (a[2] * q[0] - a[0] * q[2]) + pose.velocity[1];
    out[2] = (a[0] * q[1] - a[1] * q[0]) + pose.velocity[2];
}

struct Fragment {
    int count;        // 0 (clipped), else 1..3 stores in shader order
    int kind[3];      // 0 = the fragment's own voxel, 1 = depth - 1, 2 = depth + 1 (later kinds overwrite earlier ones)
    int cell[3][3];   // voxel written by store k
    float vel[3][3];  // its velocity
};

// conservative_hull.frag:25-52 for the fragment of pixel (px, py); the caller has checked pixel_overlaps
BLUB_HD void shade_fragment(const TriSetup &t, const MeshPose &pose, const int res[3], int viewport, int px, int py, Fragment &f) {
    f.count = 0;
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f; // gl_FragCoord.xy
    float z = (t.p[0][2] + t.gx * (fx - t.p[0][0])) + t.gy * (fy - t.p[0][1]);
    z = clampf(z, t.zmin, t.zmax);
    if (z < 0.0f || z > (float)viewport) return; // depth clip
    const float hi[3] = {(float)res[0] - 1.0f, (float)res[1] - 1.0f, (float)res[2] - 1.0f};
    const float max_change = fmaxf(fabsf(t.gx), fabsf(t.gy));
    const float sw[3][3] = {{truncf(fx), truncf(fy), truncf(z)}, {fx, fy, z - 1.0f}, {fx, fy, z + 1.0f}};
    const bool take[3] = {true, floorf(z) != floorf(z - max_change), floorf(z) != floorf(z + max_change)};
    for (int k = 0; k < 3; ++k) {
        if (!take[k]) continue;
        float pos[3];
        swizzle(t.side, sw[k], pos); // Unswizzle
        for (int d = 0; d < 3; ++d) pos[d] = clampf(pos[d], 0.0f, hi[d]);
        f.kind[f.count] = k;
        for (int d = 0; d < 3; ++d) f.cell[f.count][d] = (int)pos[d];
        voxel_speed(pose, pos, f.vel[f.count]);
        ++f.count;
    }
}

} // namespace vox
} // namespace blub
