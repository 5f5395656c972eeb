// Host twin of the mesh voxelizer's kernels, for CPU-only tests: the SAME per-triangle / per-pixel functions the device code runs
// (blub_b200/csrc/voxelize_core.hpp), driven sequentially.  Test infrastructure: built by tests/test_zz_mesh_voxelizer.py with g++,
// never part of libblubcore.so.
#include <cstdint>
#include <cstring>

#include "../../blub_b200/csrc/voxelize_core.hpp"

using namespace blub::vox;

extern "C" int voxcore_host_voxelize(const float *positions, uint32_t num_vertices, const uint32_t *indices, uint32_t num_triangles, const float *pose21,
                                     const int32_t res[3], float *rgba, int64_t *owner) {
    MeshPose pose;
    std::memcpy(pose.m, pose21, 12 * sizeof(float));
    std::memcpy(pose.centre, pose21 + 12, 3 * sizeof(float));
    std::memcpy(pose.axis, pose21 + 15, 3 * sizeof(float));
    std::memcpy(pose.velocity, pose21 + 18, 3 * sizeof(float));
    const int viewport = res[0] > res[1] ? (res[0] > res[2] ? res[0] : res[2]) : (res[1] > res[2] ? res[1] : res[2]);
    const int r[3] = {res[0], res[1], res[2]};
    for (uint32_t t = 0; t < num_triangles; ++t) {
        float v[3][3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t idx = indices[3 * t + k];
            if (idx >= num_vertices) return 1;
            transform_vertex(pose, positions + 3 * idx, v[k]);
        }
        TriSetup s;
        if (!setup_triangle(v[0], v[1], v[2], viewport, s)) continue;
        for (int py = s.y0; py <= s.y1; ++py)
            for (int px = s.x0; px <= s.x1; ++px) {
                if (!pixel_overlaps(s, px, py)) continue;
                Fragment f;
                shade_fragment(s, pose, r, viewport, px, py, f);
                for (int k = 0; k < f.count; ++k) {
                    const int64_t cell = ((int64_t)f.cell[k][2] * r[1] + f.cell[k][1]) * r[0] + f.cell[k][0];
                    const int64_t prio = ((int64_t)(4 * (int64_t)t + f.kind[k] + 1) << 32) | (int64_t)(py * viewport + px);
                    if (owner[cell] <= prio) {
                        owner[cell] = prio;
                        rgba[4 * cell + 0] = f.vel[k][0];
                        rgba[4 * cell + 1] = f.vel[k][1];
                        rgba[4 * cell + 2] = f.vel[k][2];
                        rgba[4 * cell + 3] = 1.0f;
                    }
                }
            }
    }
    return 0;
}
