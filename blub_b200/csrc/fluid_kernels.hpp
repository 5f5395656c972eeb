// fluid_kernels.hpp -- launchers of the non-PCG kernels (fluid_kernels.cu).  All enqueue-only on `st`.
#pragma once
#include "common.cuh"

namespace blub {

// np_upper: host-side upper bound of the particle count used to size the launch; the kernels guard with
// StepParams::num_particles (device), mirroring NumParticles in the reference's uniform buffer (hybrid_fluid.glsl:7-10).
// Coarse occupancy maps of the marker volume, refreshed by every boundary-marker pass (see boundary_marker_kernel).
struct MarkerFlags {
    uint8_t *seg_fluid; // n >> seg_shift entries
    uint8_t *row_fluid; // ny * nz entries
    uint8_t *row_near;  // ny * nz entries: row_fluid dilated by [-1, +2] in y and z
    int seg_shift;      // 5 (32-cell segments) or 3 when nx is not a multiple of 32
    uint8_t *face_valid; // EXPERIMENTAL (BLUB_EXTRAPOLATE=bytes), else null: per cell, bit c = "face c carries a valid velocity"
                         // (the cell or its +c neighbour is FLUID, extrapolate_velocity.comp:5-10), bit 3 = the cell is FLUID
};

void launch_p2g(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float4 *const row[3],
                float *const u[3], float2 *const nw[3], int8_t *marker, const uint2 *vox, const MarkerFlags &flags);
// the two halves of launch_p2g / launch_density_rhs, for callers that exchange halos in between (z-slab sharding)
void launch_p2g_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float4 *const row[3],
                        float2 *const nw[3], int8_t *marker);
void launch_p2g_finish(cudaStream_t st, const GridDim &g, const StepParams *params, float *const u[3], float2 *const nw[3], int8_t *marker,
                       const uint2 *vox, const MarkerFlags &flags);
void launch_density_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float *density);
void launch_density_finish(cudaStream_t st, const GridDim &g, const StepParams *params, const int8_t *marker, const float *density, float *rhs);
void launch_divergence_compute(cudaStream_t st, const GridDim &g, const int8_t *marker, float *const u[3], const uint2 *vox, float *rhs);
void launch_divergence_remove(cudaStream_t st, const GridDim &g, const int8_t *marker, const float *p, const uint2 *vox, float *const u[3]);
void launch_extrapolate(cudaStream_t st, const GridDim &g, const int8_t *marker, const MarkerFlags &flags, float *const u[3]);
void launch_clear_marker(cudaStream_t st, const GridDim &g, int8_t *marker);
void launch_advect(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                   float *const u[3], const uint2 *vox, int8_t *marker);
void launch_advect_migrate(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                           float *const u[3], const uint2 *vox, int8_t *marker, const MigrateOut &mig);
void launch_boundary_marker(cudaStream_t st, const GridDim &g, int8_t *marker, const uint2 *vox, const MarkerFlags &flags);
void launch_density_rhs(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos,
                        const int8_t *marker, float *density, float *rhs);
void launch_position_change(cudaStream_t st, const GridDim &g, const StepParams *params, const int8_t *marker, const float *p, float *const u[3]);
void launch_correct_particles(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos,
                              const int8_t *marker, float *const u[3]);
int binning_scan_blocks(const GridDim &g);
void launch_binning(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *src, float4 *dst,
                    uint32_t *cell_count, uint32_t *block_sums);

} // namespace blub
