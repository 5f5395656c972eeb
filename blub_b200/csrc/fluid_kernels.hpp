// fluid_kernels.hpp -- launchers of the non-PCG kernels (fluid_kernels.cu, transfer_kernels.cu).  All enqueue-only on `st`.
#pragma once
#include "common.cuh"

namespace blub {

// np_upper: host-side upper bound of the particle count used to size the launch; the kernels guard with
// StepParams::num_particles (device), mirroring NumParticles in the reference's uniform buffer (hybrid_fluid.glsl:7-10).

// 1 bit per cell: "the cell is FLUID", x fastest, `wpr` 32-bit words per row of nx cells (rows start on a word).  Rebuilt by every
// pass that finishes a marker volume; the extrapolation works on it.
struct FluidBits {
    uint32_t *words; // wpr * ny * nz
    int wpr;
};

// Cells that hold more than 32 particles (see cell_canonicalize_kernel): list, per-cell table slot, and the table of their 18 face sums
// (one component at a time).
struct CrowdedCells {
    uint32_t *count;        // number of crowded cells of the current lists
    uint32_t *cells;        // their cell indices, max_num_particles / 33 + 1 entries
    uint32_t *slot_of_cell; // n entries, only valid for crowded cells
    float2 *sums;           // 18 per crowded cell
};

// Per-step cell lists (counting sort of particle indices by primal cell, canonical order inside a cell): the particles of cell c are
// order[cell_start[c] .. cell_start[c + 1]).
struct CellLists {
    uint32_t *cell_start; // n + 1 entries
    uint32_t *order;      // max_num_particles entries
    uint32_t *arrival;    // scratch: the lists in arrival order of the count atomics, before canonicalisation
    uint2 *cell_slot;     // scratch: (cell, arrival slot) per particle
    uint32_t *block_sums; // scratch of the scan
    CrowdedCells crowd;
};

void configure_transfer_kernels(); // once per device, before the first launch (dynamic shared memory opt-in)
// clamp_lo: 1.0 velocity transfer, 0.5 density transfer, 0.0 binning (see transfer_position)
void launch_cell_lists(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float clamp_lo, const CellLists &l);
void launch_marker_from_lists(cudaStream_t st, const GridDim &g, const CellLists &l, int8_t *marker, const uint2 *vox, const FluidBits &bits);
void launch_p2g_gather(cudaStream_t st, const GridDim &g, const StepParams *params, const CellLists &l, const float4 *pos, float4 *const row[3],
                       const int8_t *marker, float *const u[3]);
// scatter form of P2G / density in two halves, for callers that exchange halos in between (z-slab sharding)
void launch_p2g_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float4 *const row[3],
                        float2 *const nw[3], int8_t *marker, bool clear_accumulators);
// particle_words: wpr * ny * nz words, 1 bit per cell "a particle marked this cell" (before the boundary rule); the finish pass visits and
// re-zeroes the accumulators of the one-cell dilation of those cells only
void launch_p2g_finish(cudaStream_t st, const GridDim &g, const StepParams *params, float *const u[3], float2 *const nw[3], int8_t *marker,
                       const uint2 *vox, const FluidBits &bits, uint32_t *particle_words);
void launch_density_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float *density);
void launch_density_finish(cudaStream_t st, const GridDim &g, const StepParams *params, const int8_t *marker, const float *density, float *rhs);
void launch_density_rhs(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos,
                        const int8_t *marker, float *density, float *rhs);
void launch_boundary_marker(cudaStream_t st, const GridDim &g, int8_t *marker, const uint2 *vox, const FluidBits &bits, uint32_t *particle_words = nullptr);
void launch_fluid_bits(cudaStream_t st, const GridDim &g, const int8_t *marker, const FluidBits &bits);
int binning_scan_blocks(const GridDim &g);
void launch_binning(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *src, float4 *dst, const CellLists &l);

// grid passes: they visit the cells within one cell of the fluid only (FluidBits), everything further away keeps its previous value
void launch_divergence_compute(cudaStream_t st, const GridDim &g, const FluidBits &bits, const int8_t *marker, float *const u[3], const uint2 *vox, float *rhs);
void launch_divergence_remove(cudaStream_t st, const GridDim &g, const FluidBits &bits, const int8_t *marker, const float *p, const uint2 *vox, float *const u[3]);
void launch_extrapolate(cudaStream_t st, const GridDim &g, const FluidBits &bits, float *const u[3]);
void launch_clear_marker(cudaStream_t st, const GridDim &g, int8_t *marker);
void launch_advect(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                   float *const u[3], const uint2 *vox, int8_t *marker);
void launch_advect_migrate(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                           float *const u[3], const uint2 *vox, int8_t *marker, const MigrateOut &mig);
void launch_position_change(cudaStream_t st, const GridDim &g, const FluidBits &bits, const StepParams *params, const int8_t *marker, const float *p, float *const u[3]);
void launch_correct_particles(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos,
                              const int8_t *marker, float *const u[3]);

} // namespace blub
