// fluid_kernels.cu -- every kernel of the fluid step except the pressure solve (sm_100a).
//
// Reference counterparts (relative to /root/reference/shader/simulation): transfer_clear.comp,
// transfer_build_linkedlist.comp, transfer_set_boundary_marker.comp, transfer_gather_velocity.comp,
// divergence_compute.comp, divergence_remove.comp, extrapolate_velocity.comp, advect_particles.comp,
// density_projection_{gather_error,position_change,correct_particles}.comp, particle_binning_*.comp.
//
// Not a port: the reference threads per-dual-cell linked lists through the particle buffer and gathers them with
// 729-thread groups in lock-step rounds (capped at 12 / 32 entries).  Here particle->grid transfers are scatters
// (RED.ADD.F32 into num/weight volumes, then one normalisation pass), which needs no lists, has no cap and touches each
// particle once; binning is a counting sort with a work-efficient scan and ping-pong buffers (no copy-back).
#include <cstdlib>
#include <cstring>

#include "fluid_kernels.hpp"

namespace blub {
namespace {

constexpr int PT = 256; // threads per block for particle and cell kernels

// Cell indices fit 32 bits (HybridFluid::new rejects grids of 2^31 cells or more): no 64-bit multiplies / divisions in the
// per-cell and per-particle kernels.
__device__ __forceinline__ int lin(const GridDim &g, int x, int y, int z) { return (z * g.ny + y) * g.nx + x; }
__device__ __forceinline__ void cell_of(const GridDim &g, int64_t i64, int &x, int &y, int &z) {
    const unsigned i = (unsigned)i64, nx = (unsigned)g.nx, ny = (unsigned)g.ny;
    const unsigned t = i / nx;
    x = (int)(i - t * nx);
    z = (int)(t / ny);
    y = (int)(t - (unsigned)z * ny);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
__device__ __forceinline__ float signf(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// ------------------------------------------------------------------------------------------------ P2G
// One thread per particle.  For component c the particle lies in dual cell d = trunc(pos - off_c), off_c = 0.5 except
// 1.0 on axis c (transfer_build_linkedlist.comp:21-23), and contributes to the eight faces d + {0,1}^3 -- exactly the
// set of (face, particle) pairs the reference's gather visits (transfer_gather_velocity.comp:39-97) -- with
//   weight = prod_k sat(1 - |q_k - pos_k|),  value = row_c . (q - pos, 1)          (:23-31)
// (sum w*value, sum w) of a face are interleaved as one float2 so that each contribution is a single 8-byte vector
// reduction.  MARK: also set marker[trunc(pos)] = FLUID (transfer_build_linkedlist.comp:17-19).
template <bool MARK>
__global__ void __launch_bounds__(PT) p2g_scatter_kernel(GridDim g, const StepParams *__restrict__ params,
                                                         const float4 *__restrict__ pos, const float4 *__restrict__ rowx,
                                                         const float4 *__restrict__ rowy, const float4 *__restrict__ rowz,
                                                         float2 *__restrict__ nwx, float2 *__restrict__ nwy, float2 *__restrict__ nwz,
                                                         int8_t *__restrict__ marker) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    float4 p = pos[i];
    // memory safety for ANY input: the eight faces of every dual cell must exist (a simulated particle is always inside
    // [1.001, dim - 1.001], so this only ever changes particles handed in from outside the domain)
    p.x = fminf(fmaxf(p.x, 1.0f), (float)g.nx - 1.0f);
    p.y = fminf(fmaxf(p.y, 1.0f), (float)g.ny - 1.0f);
    p.z = fminf(fmaxf(p.z, 1.0f), (float)g.nz - 1.0f);
    if (MARK) marker[lin(g, min((int)p.x, g.nx - 1), min((int)p.y, g.ny - 1), min((int)p.z, g.nz - 1))] = (int8_t)CELL_FLUID;
    const float4 rows[3] = {rowx[i], rowy[i], rowz[i]};
    float2 *const nw[3] = {nwx, nwy, nwz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ox = c == 0 ? 1.0f : 0.5f, oy = c == 1 ? 1.0f : 0.5f, oz = c == 2 ? 1.0f : 0.5f;
        const int dx = (int)(p.x - ox), dy = (int)(p.y - oy), dz = (int)(p.z - oz);
        // sample point of face (dx,dy,dz): cell + 0.5 + 0.5 e_c
        const float qx = (float)dx + ox, qy = (float)dy + oy, qz = (float)dz + oz;
        const float4 r = rows[c];
        float tx[2] = {qx - p.x, qx + 1.0f - p.x}, ty[2] = {qy - p.y, qy + 1.0f - p.y}, tz[2] = {qz - p.z, qz + 1.0f - p.z};
        float wxs[2] = {saturatef(1.0f - fabsf(tx[0])), saturatef(1.0f - fabsf(tx[1]))};
        float wys[2] = {saturatef(1.0f - fabsf(ty[0])), saturatef(1.0f - fabsf(ty[1]))};
        float wzs[2] = {saturatef(1.0f - fabsf(tz[0])), saturatef(1.0f - fabsf(tz[1]))};
        const int base = lin(g, dx, dy, dz);
#pragma unroll
        for (int oz_ = 0; oz_ < 2; ++oz_)
#pragma unroll
            for (int oy_ = 0; oy_ < 2; ++oy_)
#pragma unroll
                for (int ox_ = 0; ox_ < 2; ++ox_) {
                    const float w = wxs[ox_] * wys[oy_] * wzs[oz_];
                    if (w <= 0.0f) continue;
                    const float v = r.x * tx[ox_] + r.y * ty[oy_] + r.z * tz[oz_] + r.w;
                    const int f = base + ox_ + oy_ * g.sy + oz_ * g.sz;
                    // one 8-byte vector reduction (RED.E.ADD.F32x2) per face instead of two scalar ones.  Pairing x-adjacent
                    // faces into 16-byte F32x4 reductions when aligned was measured SLOWER (1.69 vs 1.49 ms for the stage at
                    // 16.4 M particles): the L2 atomic units are bound by sectors touched, not by instructions.
                    atomicAdd(nw[c] + f, make_float2(w * v, w));
                }
    }
}

// EXPERIMENTAL alternatives (BLUB_SCATTER=aggregate; not the default, not yet measured).  After a re-sort the particles of one cell
// sit in adjacent lanes and hit the SAME eight faces: the reductions of a warp then serialise in the L2 atomic units.  Here every run
// of adjacent lanes with the same dual cell first adds its contributions up with shuffles (segmented reduction, log2 steps) and only
// the first lane of the run issues reductions.  Same pairs, same weights; only the summation order differs.
template <int NV>
__device__ __forceinline__ bool segmented_run_sum(int key, float (&v)[NV]) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int prev = __shfl_up_sync(full, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(full, head);
    const unsigned above = lane == 31 ? 0u : (heads & ~((2u << lane) - 1u)); // run heads in the lanes above this one
    const int end = above ? __ffs(above) - 1 : 32;                            // first lane that is not part of this lane's run
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const bool take = lane + o < end;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float t = __shfl_down_sync(full, v[k], o);
            if (take) v[k] += t;
        }
    }
    return head;
}

template <bool MARK>
__global__ void __launch_bounds__(PT) p2g_scatter_aggregate_kernel(GridDim g, const StepParams *__restrict__ params, const float4 *__restrict__ pos,
                                                                   const float4 *__restrict__ rowx, const float4 *__restrict__ rowy,
                                                                   const float4 *__restrict__ rowz, float2 *__restrict__ nwx, float2 *__restrict__ nwy,
                                                                   float2 *__restrict__ nwz, int8_t *__restrict__ marker) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    const bool valid = i < params->num_particles; // no early return: every lane takes part in the shuffles
    float4 p = valid ? pos[i] : make_float4(1.5f, 1.5f, 1.5f, 0.0f);
    p.x = fminf(fmaxf(p.x, 1.0f), (float)g.nx - 1.0f);
    p.y = fminf(fmaxf(p.y, 1.0f), (float)g.ny - 1.0f);
    p.z = fminf(fmaxf(p.z, 1.0f), (float)g.nz - 1.0f);
    if (MARK && valid) marker[lin(g, min((int)p.x, g.nx - 1), min((int)p.y, g.ny - 1), min((int)p.z, g.nz - 1))] = (int8_t)CELL_FLUID;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 rows[3] = {valid ? rowx[i] : zero, valid ? rowy[i] : zero, valid ? rowz[i] : zero};
    float2 *const nw[3] = {nwx, nwy, nwz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ox = c == 0 ? 1.0f : 0.5f, oy = c == 1 ? 1.0f : 0.5f, oz = c == 2 ? 1.0f : 0.5f;
        const int dx = (int)(p.x - ox), dy = (int)(p.y - oy), dz = (int)(p.z - oz);
        const float qx = (float)dx + ox, qy = (float)dy + oy, qz = (float)dz + oz;
        const float4 r = rows[c];
        const float tx[2] = {qx - p.x, qx + 1.0f - p.x}, ty[2] = {qy - p.y, qy + 1.0f - p.y}, tz[2] = {qz - p.z, qz + 1.0f - p.z};
        const float wxs[2] = {saturatef(1.0f - fabsf(tx[0])), saturatef(1.0f - fabsf(tx[1]))};
        const float wys[2] = {saturatef(1.0f - fabsf(ty[0])), saturatef(1.0f - fabsf(ty[1]))};
        const float wzs[2] = {saturatef(1.0f - fabsf(tz[0])), saturatef(1.0f - fabsf(tz[1]))};
        const int base = lin(g, dx, dy, dz);
        float acc[16]; // (sum w * value, sum w) of the eight faces
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ox_ = k & 1, oy_ = (k >> 1) & 1, oz_ = k >> 2;
            const float w = valid ? wxs[ox_] * wys[oy_] * wzs[oz_] : 0.0f;
            const float v = r.x * tx[ox_] + r.y * ty[oy_] + r.z * tz[oz_] + r.w;
            acc[2 * k] = w > 0.0f ? w * v : 0.0f;
            acc[2 * k + 1] = w > 0.0f ? w : 0.0f;
        }
        const bool head = segmented_run_sum<16>(valid ? base : -1 - (int)(threadIdx.x & 31), acc);
        if (head && valid) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (acc[2 * k + 1] > 0.0f) atomicAdd(nw[c] + base + (k & 1) + ((k >> 1) & 1) * g.sy + (k >> 2) * g.sz, make_float2(acc[2 * k], acc[2 * k + 1]));
        }
    }
}

__global__ void __launch_bounds__(PT) density_scatter_aggregate_kernel(GridDim g, const StepParams *__restrict__ params, const float4 *__restrict__ pos,
                                                                       float *__restrict__ density) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    const bool valid = i < params->num_particles;
    float4 p = valid ? pos[i] : make_float4(1.5f, 1.5f, 1.5f, 0.0f);
    p.x = fminf(fmaxf(p.x, 0.5f), (float)g.nx - 1.0f);
    p.y = fminf(fmaxf(p.y, 0.5f), (float)g.ny - 1.0f);
    p.z = fminf(fmaxf(p.z, 0.5f), (float)g.nz - 1.0f);
    const int dx = (int)(p.x - 0.5f), dy = (int)(p.y - 0.5f), dz = (int)(p.z - 0.5f);
    const float qx = (float)dx + 0.5f, qy = (float)dy + 0.5f, qz = (float)dz + 0.5f;
    const float wx[2] = {saturatef(1.0f - fabsf(qx - p.x)), saturatef(1.0f - fabsf(qx + 1.0f - p.x))};
    const float wy[2] = {saturatef(1.0f - fabsf(qy - p.y)), saturatef(1.0f - fabsf(qy + 1.0f - p.y))};
    const float wz[2] = {saturatef(1.0f - fabsf(qz - p.z)), saturatef(1.0f - fabsf(qz + 1.0f - p.z))};
    const int base = lin(g, dx, dy, dz);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = valid ? wx[k & 1] * wy[(k >> 1) & 1] * wz[k >> 2] : 0.0f;
    const bool head = segmented_run_sum<8>(valid ? base : -1 - (int)(threadIdx.x & 31), acc);
    if (head && valid) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (acc[k] > 0.0f) atomicAdd(density + base + (k & 1) + ((k >> 1) & 1) * g.sy + (k >> 2) * g.sz, acc[k]);
    }
}

// transfer_set_boundary_marker.comp:11-20.  Also publishes two coarse occupancy maps of the finished marker volume,
// used by the extrapolation pass to reject cells that have no FLUID cell anywhere near them:
//   seg_fluid[cell / seg_w]  any FLUID cell in the seg_w x-consecutive cells (seg_w = 32, or 8 when nx % 32 != 0)
//   row_fluid[z * ny + y]    any FLUID cell in the row (must be zeroed before the launch)
__global__ void __launch_bounds__(PT) boundary_marker_kernel(GridDim g, int8_t *__restrict__ marker, const uint2 *__restrict__ vox,
                                                             uint8_t *__restrict__ seg_fluid, uint8_t *__restrict__ row_fluid, int seg_w) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; // g.n is a multiple of 512: no partial warps
    int x, y, z;
    cell_of(g, i, x, y, z);
    int m = marker[i];
    if (x == 0 || y == 0 || z <= g.z_wall_lo || x == g.nx - 1 || y == g.ny - 1 || z >= g.z_wall_hi) {
        m = CELL_SOLID;
        marker[i] = (int8_t)m;
    } else if (vox != nullptr) {
        if (load_voxel(vox, i).w != 0.0f) {
            m = CELL_SOLID;
            marker[i] = (int8_t)m;
        }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, m == CELL_FLUID);
    const int lane = threadIdx.x & 31;
    if (seg_w == 32) {
        if (lane == 0) {
            seg_fluid[i >> 5] = ballot ? 1 : 0;
            if (ballot) row_fluid[z * g.ny + y] = 1;
        }
    } else if ((lane & 7) == 0) {
        const unsigned b = (ballot >> lane) & 0xffu;
        seg_fluid[i >> 3] = b ? 1 : 0;
        if (b) row_fluid[z * g.ny + y] = 1;
    }
}

// row_near[z * ny + y] = any FLUID cell in rows [y-1, y+2] x [z-1, z+2]: one byte answers "nothing to extrapolate here"
__global__ void __launch_bounds__(PT) row_near_kernel(GridDim g, const uint8_t *__restrict__ row_fluid, uint8_t *__restrict__ row_near) {
    const int r = blockIdx.x * PT + threadIdx.x;
    if (r >= g.ny * g.nz) return;
    const int y = r % g.ny, z = r / g.ny;
    const int y0 = max(y - 1, 0), y1 = min(y + 2, g.ny - 1), z0 = max(z - 1, 0), z1 = min(z + 2, g.nz - 1);
    unsigned near = 0;
    for (int zz = z0; zz <= z1; ++zz)
        for (int yy = y0; yy <= y1; ++yy) near |= row_fluid[zz * g.ny + yy];
    row_near[r] = near ? 1 : 0;
}

// Normalisation + global forces + "don't flow into solid": transfer_gather_velocity.comp:116-127.
// Faces that touch no FLUID cell are written 0 here (the reference leaves them stale; never observable).
__global__ void __launch_bounds__(PT) p2g_normalize_kernel(GridDim g, const StepParams *__restrict__ params,
                                                           const int8_t *__restrict__ marker, float *__restrict__ ux,
                                                           float *__restrict__ uy, float *__restrict__ uz,
                                                           const float2 *__restrict__ nwx, const float2 *__restrict__ nwy,
                                                           const float2 *__restrict__ nwz) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    const int ma = marker[i];
    const int mb[3] = {marker[i + 1], marker[i + g.sy], marker[i + g.sz]};
    float *const u[3] = {ux, uy, uz};
    const float2 *const nw[3] = {nwx, nwy, nwz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float out = 0.0f;
        if (ma == CELL_FLUID || mb[c] == CELL_FLUID) {
            if (ma != CELL_SOLID && mb[c] != CELL_SOLID) {
                const float2 a = nw[c][i];
                float v = a.x;
                if (a.y > 0.0f) v /= a.y;
                out = v + params->gravity_dt[c];
            }
        }
        u[c][i] = out;
    }
}

// divergence_compute.comp:28-86
__global__ void __launch_bounds__(PT) divergence_compute_kernel(GridDim g, const int8_t *__restrict__ marker,
                                                                const float *__restrict__ ux, const float *__restrict__ uy,
                                                                const float *__restrict__ uz, const uint2 *__restrict__ vox,
                                                                float *__restrict__ rhs) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    if (marker[i] != CELL_FLUID) return;
    const float px = ux[i], py = uy[i], pz = uz[i];
    const float nx = ux[i - 1], ny = uy[i - g.sy], nz = uz[i - g.sz];
    float d = px - nx;
    d += py - ny;
    d += pz - nz;
    if (marker[i - 1] == CELL_SOLID) d += nx - load_voxel(vox, i - 1).x;
    if (marker[i - g.sy] == CELL_SOLID) d += ny - load_voxel(vox, i - g.sy).y;
    if (marker[i - g.sz] == CELL_SOLID) d += nz - load_voxel(vox, i - g.sz).z;
    if (marker[i + 1] == CELL_SOLID) d -= px - load_voxel(vox, i + 1).x;
    if (marker[i + g.sy] == CELL_SOLID) d -= py - load_voxel(vox, i + g.sy).y;
    if (marker[i + g.sz] == CELL_SOLID) d -= pz - load_voxel(vox, i + g.sz).z;
    rhs[i] = d;
}

// divergence_remove.comp:19-49
__global__ void __launch_bounds__(PT) divergence_remove_kernel(GridDim g, const int8_t *__restrict__ marker,
                                                               const float *__restrict__ p, const uint2 *__restrict__ vox,
                                                               float *__restrict__ ux, float *__restrict__ uy, float *__restrict__ uz) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    int x, y, z;
    cell_of(g, i, x, y, z);
    const int mc = marker[i];
    const float pc = mc == CELL_FLUID ? p[i] : 0.0f;
    float *const u[3] = {ux, uy, uz};
    const int64_t nb[3] = {i + 1, i + g.sy, i + g.sz};
    const bool inb[3] = {x + 1 < g.nx, y + 1 < g.ny, z + 1 < g.nz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int mn = inb[c] ? marker[nb[c]] : CELL_SOLID;
        float v = 0.0f;
        if (mc == CELL_FLUID || mn == CELL_FLUID) {
            if (mc == CELL_SOLID) {
                const Voxel s = load_voxel(vox, i);
                v = c == 0 ? s.x : (c == 1 ? s.y : s.z);
            } else if (mn == CELL_SOLID) {
                Voxel s = {0, 0, 0, 0};
                if (inb[c]) s = load_voxel(vox, nb[c]);
                v = c == 0 ? s.x : (c == 1 ? s.y : s.z);
            } else {
                const float pn = mn == CELL_FLUID ? p[nb[c]] : 0.0f;
                v = u[c][i] - (pc - pn);
            }
        }
        u[c][i] = v;
    }
}

// density_projection_position_change.comp:18-51 (writes the displacement field INTO the velocity volumes)
__global__ void __launch_bounds__(PT) position_change_kernel(GridDim g, const StepParams *__restrict__ params,
                                                             const int8_t *__restrict__ marker, const float *__restrict__ p,
                                                             float *__restrict__ ux, float *__restrict__ uy, float *__restrict__ uz) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    int x, y, z;
    cell_of(g, i, x, y, z);
    const float dt = params->dt;
    const int mc = marker[i];
    const float pc = mc == CELL_FLUID ? p[i] : 0.0f;
    float *const u[3] = {ux, uy, uz};
    const int64_t nb[3] = {i + 1, i + g.sy, i + g.sz};
    const bool inb[3] = {x + 1 < g.nx, y + 1 < g.ny, z + 1 < g.nz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int mn = inb[c] ? marker[nb[c]] : CELL_SOLID;
        const float pn = mn == CELL_FLUID ? p[nb[c]] : 0.0f;
        float d = (pn - pc) * dt;
        if (mc == CELL_SOLID || mn == CELL_SOLID) d = 0.0f;
        u[c][i] = d;
    }
}

// extrapolate_velocity.comp:26-90.  In place: only invalid faces are written, only valid faces are read.
__device__ __forceinline__ bool valid_velocity(const GridDim &g, const int8_t *__restrict__ marker, int x, int y, int z, int c) {
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) return false;
    const int64_t i = lin(g, x, y, z);
    if (marker[i] == CELL_FLUID) return true;
    const int64_t n = i + (c == 0 ? 1 : (c == 1 ? g.sy : g.sz));
    const bool in = c == 0 ? x + 1 < g.nx : (c == 1 ? y + 1 < g.ny : z + 1 < g.nz);
    return in && marker[n] == CELL_FLUID;
}
__global__ void __launch_bounds__(PT) extrapolate_kernel(GridDim g, const int8_t *__restrict__ marker, const uint8_t *__restrict__ seg_fluid,
                                                         const uint8_t *__restrict__ row_fluid, const uint8_t *__restrict__ row_near, int seg_shift,
                                                         float *__restrict__ ux,
                                                         float *__restrict__ uy, float *__restrict__ uz) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    if (marker[i] == CELL_FLUID) return;
    int x, y, z;
    cell_of(g, i, x, y, z);
    // Quick reject: a face is only written if a FLUID cell lies in [x-1,x+2] x [y-1,y+2] x [z-1,z+2] (the in-plane ring of
    // candidate faces plus their +e_c neighbours).  Rows first (warp-uniform loads), then x-segments.
    {
        if (!row_near[z * g.ny + y]) return;
        const int y0 = max(y - 1, 0), y1 = min(y + 2, g.ny - 1), z0 = max(z - 1, 0), z1 = min(z + 2, g.nz - 1);
        bool near = false;
        const int s0 = max(x - 1, 0) >> seg_shift, s1 = min(x + 2, g.nx - 1) >> seg_shift, segs = g.nx >> seg_shift;
        for (int zz = z0; zz <= z1; ++zz)
            for (int yy = y0; yy <= y1; ++yy) {
                if (!row_fluid[zz * g.ny + yy]) continue;
                const uint8_t *row = seg_fluid + (int64_t)(zz * g.ny + yy) * segs;
                for (int ss = s0; ss <= s1; ++ss) near = near || row[ss];
            }
        if (!near) return;
    }
    float *const u[3] = {ux, uy, uz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const bool in = c == 0 ? x + 1 < g.nx : (c == 1 ? y + 1 < g.ny : z + 1 < g.nz);
        const int64_t n = i + (c == 0 ? 1 : (c == 1 ? g.sy : g.sz));
        if (in && marker[n] == CELL_FLUID) continue;
        // in-plane axes, first one fastest (order of the velocityContribution lists in the shader)
        const int a = c == 0 ? 1 : 0, b = c == 2 ? 1 : 2;
        float numv = 0.0f, avg = 0.0f;
        for (int ob = -1; ob <= 1; ++ob)
            for (int oa = -1; oa <= 1; ++oa) {
                if (oa == 0 && ob == 0) continue;
                int h[3] = {x, y, z};
                h[a] += oa;
                h[b] += ob;
                if (valid_velocity(g, marker, h[0], h[1], h[2], c)) {
                    numv += 1.0f;
                    avg += u[c][lin(g, h[0], h[1], h[2])];
                }
            }
        if (numv > 0.0f) u[c][i] = avg / numv;
    }
}

// EXPERIMENTAL alternative (BLUB_EXTRAPOLATE=bytes; not the default, not yet measured): the same pass with the validity test of every
// face precomputed into one byte per cell by face_valid_kernel.  A thread then reads 18 neighbour bytes instead of 48 markers, and
// velocities only where a neighbour is valid.  Same neighbours, same order, same arithmetic: bit-identical to extrapolate_kernel.
__global__ void __launch_bounds__(PT) face_valid_kernel(GridDim g, const int8_t *__restrict__ marker, uint8_t *__restrict__ face_valid) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    int x, y, z;
    cell_of(g, i, x, y, z);
    const bool f = marker[i] == CELL_FLUID;
    const bool fx = x + 1 < g.nx && marker[i + 1] == CELL_FLUID, fy = y + 1 < g.ny && marker[i + g.sy] == CELL_FLUID,
               fz = z + 1 < g.nz && marker[i + g.sz] == CELL_FLUID;
    face_valid[i] = (uint8_t)((f || fx ? 1 : 0) | (f || fy ? 2 : 0) | (f || fz ? 4 : 0) | (f ? 8 : 0));
}
__global__ void __launch_bounds__(PT) extrapolate_bytes_kernel(GridDim g, const uint8_t *__restrict__ face_valid, const uint8_t *__restrict__ row_near,
                                                               float *__restrict__ ux, float *__restrict__ uy, float *__restrict__ uz) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    const unsigned own = face_valid[i];
    if ((own & 7u) == 7u) return; // FLUID cell, or all three faces valid already
    int x, y, z;
    cell_of(g, i, x, y, z);
    if (!row_near[z * g.ny + y]) return;
    float *const u[3] = {ux, uy, uz};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (own & (1u << c)) continue;
        const int a = c == 0 ? 1 : 0, b = c == 2 ? 1 : 2; // in-plane axes, first one fastest
        float numv = 0.0f, avg = 0.0f;
        for (int ob = -1; ob <= 1; ++ob)
            for (int oa = -1; oa <= 1; ++oa) {
                if (oa == 0 && ob == 0) continue;
                int h[3] = {x, y, z};
                h[a] += oa;
                h[b] += ob;
                if (h[0] < 0 || h[1] < 0 || h[2] < 0 || h[0] >= g.nx || h[1] >= g.ny || h[2] >= g.nz) continue;
                const int j = lin(g, h[0], h[1], h[2]);
                if (face_valid[j] & (1u << c)) {
                    numv += 1.0f;
                    avg += u[c][j];
                }
            }
        if (numv > 0.0f) u[c][i] = avg / numv;
    }
}

// ------------------------------------------------------------------------------------------------ G2P + advection
__device__ __forceinline__ Voxel voxel_point_clamp(const GridDim &g, const uint2 *__restrict__ vox, float px, float py, float pz) {
    // texture(sampler3D(SceneVoxelization, SamplerPointClamp), pos / gridSize): nearest texel, clamp to edge
    const int x = clampi((int)floorf(px), 0, g.nx - 1), y = clampi((int)floorf(py), 0, g.ny - 1), z = clampi((int)floorf(pz), 0, g.nz - 1);
    return load_voxel(vox, lin(g, x, y, z));
}
__device__ __forceinline__ float voxel_w_trilinear(const GridDim &g, const uint2 *__restrict__ vox, float ux, float uy, float uz) {
    const float cx = ux - 0.5f, cy = uy - 0.5f, cz = uz - 0.5f;
    const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    const float tx = cx - fx, ty = cy - fy, tz = cz - fz;
    const int x0 = clampi((int)fx, 0, g.nx - 1), x1 = clampi((int)fx + 1, 0, g.nx - 1);
    const int y0 = clampi((int)fy, 0, g.ny - 1), y1 = clampi((int)fy + 1, 0, g.ny - 1);
    const int z0 = clampi((int)fz, 0, g.nz - 1), z1 = clampi((int)fz + 1, 0, g.nz - 1);
    const float c00 = mixf(load_voxel(vox, lin(g, x0, y0, z0)).w, load_voxel(vox, lin(g, x1, y0, z0)).w, tx);
    const float c10 = mixf(load_voxel(vox, lin(g, x0, y1, z0)).w, load_voxel(vox, lin(g, x1, y1, z0)).w, tx);
    const float c01 = mixf(load_voxel(vox, lin(g, x0, y0, z1)).w, load_voxel(vox, lin(g, x1, y0, z1)).w, tx);
    const float c11 = mixf(load_voxel(vox, lin(g, x0, y1, z1)).w, load_voxel(vox, lin(g, x1, y1, z1)).w, tx);
    return mixf(mixf(c00, c10, ty), mixf(c01, c11, ty), tz);
}
__device__ __forceinline__ float grid_trilinear_clamp(const GridDim &g, const float *__restrict__ vol, float ux, float uy, float uz) {
    const float cx = ux - 0.5f, cy = uy - 0.5f, cz = uz - 0.5f;
    const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    const float tx = cx - fx, ty = cy - fy, tz = cz - fz;
    const int x0 = clampi((int)fx, 0, g.nx - 1), x1 = clampi((int)fx + 1, 0, g.nx - 1);
    const int y0 = clampi((int)fy, 0, g.ny - 1), y1 = clampi((int)fy + 1, 0, g.ny - 1);
    const int z0 = clampi((int)fz, 0, g.nz - 1), z1 = clampi((int)fz + 1, 0, g.nz - 1);
    const float c00 = mixf(vol[lin(g, x0, y0, z0)], vol[lin(g, x1, y0, z0)], tx);
    const float c10 = mixf(vol[lin(g, x0, y1, z0)], vol[lin(g, x1, y1, z0)], tx);
    const float c01 = mixf(vol[lin(g, x0, y0, z1)], vol[lin(g, x1, y0, z1)], tx);
    const float c11 = mixf(vol[lin(g, x0, y1, z1)], vol[lin(g, x1, y1, z1)], tx);
    return mixf(mixf(c00, c10, ty), mixf(c01, c11, ty), tz);
}

struct Corners {
    float v[8][3]; // corner order 000,100,010,110,001,101,011,111; [..][component]
};
__device__ __forceinline__ void trilerp3(const Corners &c, const float ix[3], const float iy[3], const float iz[3], float out[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { // InterpolateTrilinear, advect_particles.comp:19-23
        const float a = mixf(mixf(c.v[0][k], c.v[1][k], ix[k]), mixf(c.v[2][k], c.v[3][k], ix[k]), iy[k]);
        const float b = mixf(mixf(c.v[4][k], c.v[5][k], ix[k]), mixf(c.v[6][k], c.v[7][k], ix[k]), iy[k]);
        out[k] = mixf(a, b, iz[k]);
    }
}

// advect_particles.comp:35-194.  Writes position + the three APIC rows, marks the new cell FLUID (:175-178); the
// linked-list rebuild of :179-181 has no counterpart (the density pass scatters).
// MIGRATE (z-slab ranks): instead of writing back in place, the kernel itself sorts its results -- stayers are compacted
// into the spare arrays, particles that left the slab go straight to the neighbour (see MigrateOut) -- so that migration
// costs one warp-aggregated atomic per warp instead of an extra pass over all particles.
template <bool MIGRATE>
__global__ void __launch_bounds__(PT) advect_kernel(GridDim g, const StepParams *__restrict__ params, float4 *__restrict__ pos,
                                                    float4 *__restrict__ rowx, float4 *__restrict__ rowy, float4 *__restrict__ rowz,
                                                    const float *__restrict__ ux, const float *__restrict__ uy,
                                                    const float *__restrict__ uz, const uint2 *__restrict__ vox,
                                                    int8_t *__restrict__ marker, MigrateOut mig) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    const float dt = params->dt;
    const float4 p4 = pos[i];
    float x0[3] = {p4.x, p4.y, p4.z};
    const int S[3] = {g.nx, g.ny, g.nz};
    if (vox != nullptr) { // :45-64 particle "eaten" by a moving wall
        const Voxel s = voxel_point_clamp(g, vox, x0[0], x0[1], x0[2]);
        if (s.w > 0.0f) {
            const float ax = fabsf(s.x), ay = fabsf(s.y), az = fabsf(s.z);
            if (ax > ay) {
                if (ax > az) x0[0] += signf(s.x); else x0[2] += signf(s.z);
            } else {
                if (ay > az) x0[1] += signf(s.y); else x0[2] += signf(s.z);
            }
        }
    }
    Corners cn;
    float ix[3], iy[3], iz[3];
    const float *const U[3] = {ux, uy, uz};
#pragma unroll
    for (int c = 0; c < 3; ++c) { // :73-92
        float op[3];
        int lo[3], hi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            op[k] = fmaxf(0.0f, x0[k] - (k == c ? 1.0f : 0.5f));
            lo[k] = (int)op[k];
            hi[k] = min(lo[k] + 1, S[k] - 1);
            lo[k] = min(lo[k], S[k] - 1); // only reachable for particles outside the domain
        }
        const float *u = U[c];
        cn.v[0][c] = u[lin(g, lo[0], lo[1], lo[2])]; cn.v[1][c] = u[lin(g, hi[0], lo[1], lo[2])];
        cn.v[2][c] = u[lin(g, lo[0], hi[1], lo[2])]; cn.v[3][c] = u[lin(g, hi[0], hi[1], lo[2])];
        cn.v[4][c] = u[lin(g, lo[0], lo[1], hi[2])]; cn.v[5][c] = u[lin(g, hi[0], lo[1], hi[2])];
        cn.v[6][c] = u[lin(g, lo[0], hi[1], hi[2])]; cn.v[7][c] = u[lin(g, hi[0], hi[1], hi[2])];
        ix[c] = fractf(op[0]); iy[c] = fractf(op[1]); iz[c] = fractf(op[2]);
    }
    float nv[3], cx[3], cy[3], cz[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { // :96-112
        const float vx00 = mixf(cn.v[0][k], cn.v[1][k], ix[k]), vx01 = mixf(cn.v[4][k], cn.v[5][k], ix[k]);
        const float vx10 = mixf(cn.v[2][k], cn.v[3][k], ix[k]), vx11 = mixf(cn.v[6][k], cn.v[7][k], ix[k]);
        const float vxy0 = mixf(vx00, vx10, iy[k]), vxy1 = mixf(vx01, vx11, iy[k]);
        nv[k] = mixf(vxy0, vxy1, iz[k]);
        cx[k] = mixf(mixf(cn.v[1][k], cn.v[3][k], iy[k]), mixf(cn.v[5][k], cn.v[7][k], iy[k]), iz[k]) -
                mixf(mixf(cn.v[0][k], cn.v[2][k], iy[k]), mixf(cn.v[4][k], cn.v[6][k], iy[k]), iz[k]);
        cy[k] = mixf(vx10, vx11, iz[k]) - mixf(vx00, vx01, iz[k]);
        cz[k] = vxy1 - vxy0;
    }
    // RK4 confined to the sampled cell, :116-126.  As written (quirk B17): the shader adds the step VECTOR element-wise to vec3s that are
    // indexed by the velocity component, so component k is re-sampled with all three of its interpolants advanced by step[k].
    float k2[3], k3[3], k4[3], ax[3], ay[3], az[3], st[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) st[k] = dt * 0.5f * nv[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ax[k] = saturatef(ix[k] + st[k]); ay[k] = saturatef(iy[k] + st[k]); az[k] = saturatef(iz[k] + st[k]); }
    trilerp3(cn, ax, ay, az, k2);
#pragma unroll
    for (int k = 0; k < 3; ++k) st[k] = dt * 0.5f * k2[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ax[k] = saturatef(ix[k] + st[k]); ay[k] = saturatef(iy[k] + st[k]); az[k] = saturatef(iz[k] + st[k]); }
    trilerp3(cn, ax, ay, az, k3);
#pragma unroll
    for (int k = 0; k < 3; ++k) st[k] = dt * k3[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ax[k] = saturatef(ix[k] + st[k]); ay[k] = saturatef(iy[k] + st[k]); az[k] = saturatef(iz[k] + st[k]); }
    trilerp3(cn, ax, ay, az, k4);
    float mv[3], x1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mv[k] = dt * (1.0f / 6.0f) * (nv[k] + 2.0f * (k2[k] + k3[k]) + k4[k]);
        x1[k] = x0[k] + mv[k];
    }
    const float lo[3] = {1.001f, 1.001f, (float)g.z_wall_lo + 1.001f};
    const float hi[3] = {(float)g.nx - 1.001f, (float)g.ny - 1.001f, (float)(g.z_wall_hi + 1) - 1.001f};
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) hit = hit || (fminf(fmaxf(x1[k], lo[k]), hi[k]) != x1[k]);
    if (!hit && vox != nullptr) hit = voxel_point_clamp(g, vox, x1[0], x1[1], x1[2]).w > 0.0f;
    if (hit) { // :134-173
        const float len = sqrtf(mv[0] * mv[0] + mv[1] * mv[1] + mv[2] * mv[2]) + 1e-10f;
        const float dir[3] = {mv[0] / len, mv[1] / len, mv[2] / len};
        float maxstep = len;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pc = fractf(x0[k]);
            maxstep = fminf(maxstep, (dir[k] > 0.0f ? pc : 1.0f - pc) / fabsf(dir[k]) - 0.001f);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) mv[k] = dir[k] * maxstep;
        if ((int)x0[0] == (int)x1[0] && (int)x0[1] == (int)x1[1] && (int)x0[2] == (int)x1[2] && vox != nullptr) {
            float push[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float a[3] = {x1[0], x1[1], x1[2]}, b[3] = {x1[0], x1[1], x1[2]};
                a[k] -= 1.0f;
                b[k] += 1.0f;
                push[k] = voxel_w_trilinear(g, vox, a[0], a[1], a[2]) - voxel_w_trilinear(g, vox, b[0], b[1], b[2]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) mv[k] += push[k] * (dt * 50.0f);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x1[k] = fminf(fmaxf(x0[k] + mv[k], lo[k]), hi[k]);
            nv[k] = (dir[k] * maxstep) / dt;
        }
    }
    marker[lin(g, clampi((int)x1[0], 0, g.nx - 1), clampi((int)x1[1], 0, g.ny - 1), clampi((int)x1[2], 0, g.nz - 1))] = (int8_t)CELL_FLUID;
    const float4 out_pos = make_float4(x1[0], x1[1], x1[2], p4.w);
    const float4 out_rx = make_float4(cx[0], cx[1], cx[2], nv[0]); // :184-188 (B4: Jacobian columns stored as the rows)
    const float4 out_ry = make_float4(cy[0], cy[1], cy[2], nv[1]);
    const float4 out_rz = make_float4(cz[0], cz[1], cz[2], nv[2]);
    if (!MIGRATE) {
        pos[i] = out_pos;
        rowx[i] = out_rx;
        rowy[i] = out_ry;
        rowz[i] = out_rz;
        return;
    }
    const int dest = (x1[2] < mig.z_lo && mig.peer_down) ? 1 : ((x1[2] >= mig.z_hi && mig.peer_up) ? 2 : 0);
    const unsigned active = __activemask();
    unsigned slot = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) { // one atomic per destination per warp
        const unsigned m = __ballot_sync(active, dest == d);
        if (m == 0u) continue;
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(mig.counters + d, (unsigned)__popc(m));
        base = __shfl_sync(active, base, leader);
        if (dest == d) slot = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    }
    if (dest == 0) {
        mig.pos[slot] = out_pos;
        mig.rx[slot] = out_rx;
        mig.ry[slot] = out_ry;
        mig.rz[slot] = out_rz;
    } else if (slot < mig.capacity) {
        MigrantRecord rec = {out_pos, out_rx, out_ry, out_rz};
        rec.pos.z += dest == 1 ? mig.zshift : -mig.zshift; // re-base into the neighbour's local frame
        (dest == 1 ? mig.peer_down : mig.peer_up)[slot] = rec; // P2P store over NVLink
    } else {
        mig.counters[3] = 1u; // cannot happen with a CFL-limited flow; never overrun the neighbour's buffer
    }
}

// ------------------------------------------------------------------------------------------------ density projection
// Scatter counterpart of density_projection_gather_error.comp:41-97: dual cell d = trunc(pos - 0.5), cell centres d + {0,1}^3.
__global__ void __launch_bounds__(PT) density_scatter_kernel(GridDim g, const StepParams *__restrict__ params,
                                                             const float4 *__restrict__ pos, float *__restrict__ density) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    float4 p = pos[i];
    p.x = fminf(fmaxf(p.x, 0.5f), (float)g.nx - 1.0f); // memory safety, see p2g_scatter_kernel
    p.y = fminf(fmaxf(p.y, 0.5f), (float)g.ny - 1.0f);
    p.z = fminf(fmaxf(p.z, 0.5f), (float)g.nz - 1.0f);
    const int dx = (int)(p.x - 0.5f), dy = (int)(p.y - 0.5f), dz = (int)(p.z - 0.5f);
    const float qx = (float)dx + 0.5f, qy = (float)dy + 0.5f, qz = (float)dz + 0.5f;
    const float wx[2] = {saturatef(1.0f - fabsf(qx - p.x)), saturatef(1.0f - fabsf(qx + 1.0f - p.x))};
    const float wy[2] = {saturatef(1.0f - fabsf(qy - p.y)), saturatef(1.0f - fabsf(qy + 1.0f - p.y))};
    const float wz[2] = {saturatef(1.0f - fabsf(qz - p.z)), saturatef(1.0f - fabsf(qz + 1.0f - p.z))};
    const int base = lin(g, dx, dy, dz);
#pragma unroll
    for (int oz = 0; oz < 2; ++oz)
#pragma unroll
        for (int oy = 0; oy < 2; ++oy)
#pragma unroll
            for (int ox = 0; ox < 2; ++ox) {
                const float w = wx[ox] * wy[oy] * wz[oz];
                if (w > 0.0f) atomicAdd(density + base + ox + oy * g.sy + oz * g.sz, w);
            }
}

// density_projection_gather_error.comp:99-199
__global__ void __launch_bounds__(PT) density_rhs_kernel(GridDim g, const StepParams *__restrict__ params,
                                                         const int8_t *__restrict__ marker, const float *__restrict__ density,
                                                         float *__restrict__ rhs) {
    const int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x;
    if (i >= g.n) return;
    if (marker[i] != CELL_FLUID) return;
    float d = density[i];
    const int m[6] = {marker[i + 1], marker[i + g.sy], marker[i + g.sz], marker[i - 1], marker[i - g.sy], marker[i - g.sz]};
    bool any_air = false;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (m[k] == CELL_SOLID) d += 0.5625f;
        any_air = any_air || (m[k] == CELL_AIR);
    }
    if (any_air) d = fmaxf(8.0f, d);
    d = 1.0f - d / 8.0f;
    d = fminf(fmaxf(d, -0.5f), 0.5f);
    d /= params->dt;
    rhs[i] = d;
}

// density_projection_correct_particles.comp:25-73 (fp32 software trilinear instead of the 8-bit hardware filter)
__global__ void __launch_bounds__(PT) correct_particles_kernel(GridDim g, const StepParams *__restrict__ params,
                                                               float4 *__restrict__ pos, const int8_t *__restrict__ marker,
                                                               const float *__restrict__ ux, const float *__restrict__ uy,
                                                               const float *__restrict__ uz) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    const float4 p4 = pos[i];
    const float x0[3] = {p4.x, p4.y, p4.z};
    float ch[3];
    ch[0] = grid_trilinear_clamp(g, ux, fmaxf(0.0f, x0[0] - 0.5f), fmaxf(0.0f, x0[1]), fmaxf(0.0f, x0[2]));
    ch[1] = grid_trilinear_clamp(g, uy, fmaxf(0.0f, x0[0]), fmaxf(0.0f, x0[1] - 0.5f), fmaxf(0.0f, x0[2]));
    ch[2] = grid_trilinear_clamp(g, uz, fmaxf(0.0f, x0[0]), fmaxf(0.0f, x0[1]), fmaxf(0.0f, x0[2] - 0.5f));
    float x1[3] = {x0[0] + ch[0], x0[1] + ch[1], x0[2] + ch[2]};
    const float lo[3] = {1.001f, 1.001f, (float)g.z_wall_lo + 1.001f};
    const float hi[3] = {(float)g.nx - 1.001f, (float)g.ny - 1.001f, (float)(g.z_wall_hi + 1) - 1.001f};
    bool hit = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) hit = hit || (fminf(fmaxf(x1[k], lo[k]), hi[k]) != x1[k]);
    if (!hit) {
        const int x = clampi((int)floorf(x1[0]), 0, g.nx - 1), y = clampi((int)floorf(x1[1]), 0, g.ny - 1), z = clampi((int)floorf(x1[2]), 0, g.nz - 1);
        hit = marker[lin(g, x, y, z)] == CELL_SOLID;
    }
    if (hit) {
        const float len = sqrtf(ch[0] * ch[0] + ch[1] * ch[1] + ch[2] * ch[2]) + 1e-10f;
        const float dir[3] = {ch[0] / len, ch[1] / len, ch[2] / len};
        float maxstep = len;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pc = fractf(x0[k]);
            maxstep = fminf(maxstep, (dir[k] > 0.0f ? pc : 1.0f - pc) / fabsf(dir[k]) - 0.001f);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) x1[k] = fminf(fmaxf(x0[k] + dir[k] * maxstep, lo[k]), hi[k]);
    }
    x1[2] = fminf(fmaxf(x1[2], g.z_keep_lo), g.z_keep_hi); // slab ranks only (no-op on one GPU)
    pos[i] = make_float4(x1[0], x1[1], x1[2], p4.w);
}

// ------------------------------------------------------------------------------------------------ binning
// particle_binning_count.comp (guarded): rank inside the cell kept in pos.w
__global__ void __launch_bounds__(PT) binning_count_kernel(GridDim g, const StepParams *__restrict__ params, float4 *__restrict__ pos,
                                                           uint32_t *__restrict__ count) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    float4 p = pos[i];
    const int x = clampi((int)p.x, 0, g.nx - 1), y = clampi((int)p.y, 0, g.ny - 1), z = clampi((int)p.z, 0, g.nz - 1);
    const uint32_t rank = atomicAdd(count + lin(g, x, y, z), 1u);
    pos[i].w = __uint_as_float(rank);
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8; // 2048 cells per block
// phase 1: per-block totals
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint32_t *__restrict__ in, int64_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) acc += in[base + k];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) t += sh[k];
        sums[blockIdx.x] = t;
    }
}
// phase 2: exclusive scan of the block totals by one block (deterministic block order, unlike the reference's atomic
// arrival order, particle_binning_prefixsum.comp:53-56)
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *__restrict__ sums, int nblocks) {
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int idx = base + threadIdx.x;
        const uint32_t v = idx < nblocks ? sums[idx] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        const uint32_t incl = sh[threadIdx.x];
        if (idx < nblocks) sums[idx] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}
// phase 3: exclusive scan inside each block + block base, written in place
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(uint32_t *__restrict__ data, int64_t n, const uint32_t *__restrict__ sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? data[base + k] : 0u;
        acc += v[k];
    }
    uint32_t incl = acc;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (int k = 0; k < w; ++k) wbase += sh[k];
    uint32_t run = sums[blockIdx.x] + wbase + incl - acc;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}
// particle_binning_rewrite_particles.comp with exclusive offsets (the as-written inclusive - index is off by one, B2)
__global__ void __launch_bounds__(PT) binning_scatter_kernel(GridDim g, const StepParams *__restrict__ params,
                                                             const float4 *__restrict__ src, float4 *__restrict__ dst,
                                                             const uint32_t *__restrict__ offsets) {
    const uint32_t i = blockIdx.x * PT + threadIdx.x;
    if (i >= params->num_particles) return;
    const float4 p = src[i];
    const int x = clampi((int)p.x, 0, g.nx - 1), y = clampi((int)p.y, 0, g.ny - 1), z = clampi((int)p.z, 0, g.nz - 1);
    const uint32_t d = offsets[lin(g, x, y, z)] + __float_as_uint(p.w);
    dst[d] = make_float4(p.x, p.y, p.z, 0.0f);
}

inline int blocks_for(int64_t n, int per_block) { return (int)((n + per_block - 1) / per_block); }

} // namespace

// ------------------------------------------------------------------------------------------------ launchers
// opt-in experiment (see p2g_scatter_aggregate_kernel); read when a launch is issued (a captured step graph keeps the choice it was
// captured with)
static bool scatter_aggregate() {
    const char *e = std::getenv("BLUB_SCATTER");
    return e && std::strcmp(e, "aggregate") == 0;
}

static void run_boundary_marker(cudaStream_t st, const GridDim &g, int8_t *marker, const uint2 *vox, const MarkerFlags &flags) {
    BLUB_CUDA_CHECK(cudaMemsetAsync(flags.row_fluid, 0, (size_t)g.ny * g.nz, st));
    BLUB_LAUNCH(boundary_marker_kernel, blocks_for(g.n, PT), PT, 0, st, g, marker, vox, flags.seg_fluid, flags.row_fluid, 1 << flags.seg_shift);
    BLUB_LAUNCH(row_near_kernel, blocks_for((int64_t)g.ny * g.nz, PT), PT, 0, st, g, flags.row_fluid, flags.row_near);
    if (flags.face_valid) BLUB_LAUNCH(face_valid_kernel, blocks_for(g.n, PT), PT, 0, st, g, marker, flags.face_valid);
}

void launch_p2g_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float4 *const row[3],
                        float2 *const nw[3], int8_t *marker) {
    // transfer_clear.comp: marker <- AIR; the (num, weight) volumes replace the linked-list head volume
    BLUB_CUDA_CHECK(cudaMemsetAsync(marker, 0xFF, (size_t)g.n, st));
    for (int c = 0; c < 3; ++c) BLUB_CUDA_CHECK(cudaMemsetAsync(nw[c], 0, (size_t)g.n * sizeof(float2), st));
    if (np_upper == 0) return;
    if (scatter_aggregate())
        BLUB_LAUNCH(p2g_scatter_aggregate_kernel<true>, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, row[0], row[1], row[2], nw[0], nw[1], nw[2], marker);
    else
        BLUB_LAUNCH(p2g_scatter_kernel<true>, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, row[0], row[1], row[2], nw[0], nw[1], nw[2], marker);
}

void launch_p2g_finish(cudaStream_t st, const GridDim &g, const StepParams *params, float *const u[3], float2 *const nw[3], int8_t *marker,
                       const uint2 *vox, const MarkerFlags &flags) {
    run_boundary_marker(st, g, marker, vox, flags);
    BLUB_LAUNCH(p2g_normalize_kernel, blocks_for(g.n, PT), PT, 0, st, g, params, marker, u[0], u[1], u[2], nw[0], nw[1], nw[2]);
}

void launch_p2g(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float4 *const row[3],
                float *const u[3], float2 *const nw[3], int8_t *marker, const uint2 *vox, const MarkerFlags &flags) {
    launch_p2g_scatter(st, g, params, np_upper, pos, row, nw, marker);
    launch_p2g_finish(st, g, params, u, nw, marker, vox, flags);
}

void launch_divergence_compute(cudaStream_t st, const GridDim &g, const int8_t *marker, float *const u[3], const uint2 *vox, float *rhs) {
    BLUB_LAUNCH(divergence_compute_kernel, blocks_for(g.n, PT), PT, 0, st, g, marker, u[0], u[1], u[2], vox, rhs);
}

void launch_divergence_remove(cudaStream_t st, const GridDim &g, const int8_t *marker, const float *p, const uint2 *vox, float *const u[3]) {
    BLUB_LAUNCH(divergence_remove_kernel, blocks_for(g.n, PT), PT, 0, st, g, marker, p, vox, u[0], u[1], u[2]);
}

void launch_extrapolate(cudaStream_t st, const GridDim &g, const int8_t *marker, const MarkerFlags &flags, float *const u[3]) {
    if (flags.face_valid) {
        BLUB_LAUNCH(extrapolate_bytes_kernel, blocks_for(g.n, PT), PT, 0, st, g, flags.face_valid, flags.row_near, u[0], u[1], u[2]);
        return;
    }
    BLUB_LAUNCH(extrapolate_kernel, blocks_for(g.n, PT), PT, 0, st, g, marker, flags.seg_fluid, flags.row_fluid, flags.row_near, flags.seg_shift, u[0], u[1], u[2]);
}

void launch_clear_marker(cudaStream_t st, const GridDim &g, int8_t *marker) {
    BLUB_CUDA_CHECK(cudaMemsetAsync(marker, 0xFF, (size_t)g.n, st));
}

void launch_advect(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                   float *const u[3], const uint2 *vox, int8_t *marker) {
    if (np_upper == 0) return;
    BLUB_LAUNCH(advect_kernel<false>, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, row[0], row[1], row[2], u[0], u[1], u[2], vox, marker, MigrateOut{});
}

void launch_advect_migrate(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos, float4 *const row[3],
                           float *const u[3], const uint2 *vox, int8_t *marker, const MigrateOut &mig) {
    if (np_upper == 0) return;
    BLUB_LAUNCH(advect_kernel<true>, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, row[0], row[1], row[2], u[0], u[1], u[2], vox, marker, mig);
}

void launch_boundary_marker(cudaStream_t st, const GridDim &g, int8_t *marker, const uint2 *vox, const MarkerFlags &flags) {
    run_boundary_marker(st, g, marker, vox, flags);
}

void launch_density_scatter(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos, float *density) {
    BLUB_CUDA_CHECK(cudaMemsetAsync(density, 0, (size_t)g.n * sizeof(float), st));
    if (np_upper == 0) return;
    if (scatter_aggregate()) BLUB_LAUNCH(density_scatter_aggregate_kernel, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, density);
    else BLUB_LAUNCH(density_scatter_kernel, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, density);
}

void launch_density_finish(cudaStream_t st, const GridDim &g, const StepParams *params, const int8_t *marker, const float *density, float *rhs) {
    BLUB_LAUNCH(density_rhs_kernel, blocks_for(g.n, PT), PT, 0, st, g, params, marker, density, rhs);
}

void launch_density_rhs(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, const float4 *pos,
                        const int8_t *marker, float *density, float *rhs) {
    launch_density_scatter(st, g, params, np_upper, pos, density);
    launch_density_finish(st, g, params, marker, density, rhs);
}

void launch_position_change(cudaStream_t st, const GridDim &g, const StepParams *params, const int8_t *marker, const float *p, float *const u[3]) {
    BLUB_LAUNCH(position_change_kernel, blocks_for(g.n, PT), PT, 0, st, g, params, marker, p, u[0], u[1], u[2]);
}

void launch_correct_particles(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *pos,
                              const int8_t *marker, float *const u[3]) {
    if (np_upper == 0) return;
    BLUB_LAUNCH(correct_particles_kernel, blocks_for(np_upper, PT), PT, 0, st, g, params, pos, marker, u[0], u[1], u[2]);
}

int binning_scan_blocks(const GridDim &g) { return blocks_for(g.n, SCAN_THREADS * SCAN_ITEMS); }

void launch_binning(cudaStream_t st, const GridDim &g, const StepParams *params, uint32_t np_upper, float4 *src, float4 *dst,
                    uint32_t *cell_count, uint32_t *block_sums) {
    if (np_upper == 0) return;
    const int nb = binning_scan_blocks(g);
    BLUB_CUDA_CHECK(cudaMemsetAsync(cell_count, 0, (size_t)g.n * sizeof(uint32_t), st));
    BLUB_LAUNCH(binning_count_kernel, blocks_for(np_upper, PT), PT, 0, st, g, params, src, cell_count);
    BLUB_LAUNCH(scan_block_sums_kernel, nb, SCAN_THREADS, 0, st, cell_count, g.n, block_sums);
    BLUB_LAUNCH(scan_sums_kernel, 1, 1024, 0, st, block_sums, nb);
    BLUB_LAUNCH(scan_apply_kernel, nb, SCAN_THREADS, 0, st, cell_count, g.n, block_sums);
    BLUB_LAUNCH(binning_scatter_kernel, blocks_for(np_upper, PT), PT, 0, st, g, params, src, dst, cell_count);
}

} // namespace blub
