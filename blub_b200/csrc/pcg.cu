// pcg.cu -- preconditioned conjugate gradient pressure solve (sm_100a).
//
// Replaces PressureSolver::solve (src/simulation/pressure_solver.rs:591-729) and the seven compute pipelines under
// shader/simulation/pressure_solver/.  Same recurrence, same iteration schedule (error checks at i % freq == 0 and
// i == max, pressure_solver.rs:676-677), same epsilon guards (pressure_reduce.comp:73-80) -- but not the same pass
// structure: the reference records 9 dispatches + 3 two-level reductions per iteration (313 dispatches per solve,
// ~85 B/cell/iteration); here one iteration is three kernels
//     dot     : s.As                                       (pressure_apply_coeff.comp + reduce ALPHA)
//     update  : p += a s, r -= a As, z.r and max|r| fused  (pressure_update_pressure_and_residual.comp + both
//                                                           preconditioner passes + reduce BETA / MAX_ERROR)
//     search  : s = z + b s                                (pressure_update_search.comp)
// each ending in a deterministic last-block-done reduction, with the diagonal ("diag2") preconditioner evaluated on
// the fly so that z is never stored.  Convergence is a device flag that turns the remaining launches into no-ops
// (the reference zeroes its indirect-dispatch arguments instead, pressure_reduce.comp:89-92).
//
// Layout: four cells per thread along x (128-bit loads), marker as int8, fp32 vectors, padded arrays (common.cuh).
#include "blub_core.hpp"

namespace blub {

std::atomic<uint64_t> g_kernel_launches{0};

namespace {

constexpr int PCG_THREADS = 256;
constexpr float PCG_EPSILON = 1e-10f; // pressure_reduce.comp:33

struct TileMap {
    int bx, by, bz;                   // threads per tile edge (x in quads of 4 cells)
    int tiles_x, tiles_y, tiles_z;
    int qx;                           // quads per row
    int nblocks;
};

TileMap make_tilemap(const GridDim &g) {
    TileMap t;
    t.qx = g.nx / 4;
    t.bx = t.qx >= 32 ? 32 : (t.qx >= 16 ? 16 : 8);
    if (t.qx < 8) t.bx = t.qx; // nx = 8, 16, 24
    t.by = 4;
    t.bz = PCG_THREADS / (t.bx * t.by);
    if (t.bz > 8) t.bz = 8;
    t.tiles_x = (t.qx + t.bx - 1) / t.bx;
    t.tiles_y = g.ny / t.by;
    t.tiles_z = (g.nz + t.bz - 1) / t.bz;
    t.nblocks = t.tiles_x * t.tiles_y * t.tiles_z;
    return t;
}

__device__ __forceinline__ bool tile_cell(const GridDim &g, const TileMap &t, int64_t &i) {
    int tid = threadIdx.x;
    int lx = tid % t.bx, ly = (tid / t.bx) % t.by, lz = tid / (t.bx * t.by);
    int b = blockIdx.x;
    int tx = b % t.tiles_x, ty = (b / t.tiles_x) % t.tiles_y, tz = b / (t.tiles_x * t.tiles_y);
    int q = tx * t.bx + lx, y = ty * t.by + ly, z = tz * t.bz + lz;
    if (q >= t.qx || lz >= t.bz || z >= g.nz) return false;
    i = ((int64_t)z * g.ny + y) * g.nx + 4 * q;
    return true;
}

// Markers of a quad of four x-consecutive cells and of their six neighbours.
struct QuadStencil {
    bool fluid[4];
    float diag[4];       // number of non-SOLID neighbours (pressure.glsl:44-50)
    unsigned nbr[4];     // FLUID flags: bit 0 -x, 1 +x, 2 -y, 3 +y, 4 -z, 5 +z
};

__device__ __forceinline__ bool load_quad_stencil(const int8_t *__restrict__ m, int64_t i, int sy, int sz, QuadStencil &q) {
    const char4 c = *reinterpret_cast<const char4 *>(m + i);
    if (c.x != CELL_FLUID && c.y != CELL_FLUID && c.z != CELL_FLUID && c.w != CELL_FLUID) return false;
    const char4 ym = *reinterpret_cast<const char4 *>(m + i - sy), yp = *reinterpret_cast<const char4 *>(m + i + sy);
    const char4 zm = *reinterpret_cast<const char4 *>(m + i - sz), zp = *reinterpret_cast<const char4 *>(m + i + sz);
    const int cx[6] = {m[i - 1], c.x, c.y, c.z, c.w, m[i + 4]};
    const int cym[4] = {ym.x, ym.y, ym.z, ym.w}, cyp[4] = {yp.x, yp.y, yp.z, yp.w};
    const int czm[4] = {zm.x, zm.y, zm.z, zm.w}, czp[4] = {zp.x, zp.y, zp.z, zp.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q.fluid[k] = cx[k + 1] == CELL_FLUID;
        int d = (cx[k] != CELL_SOLID) + (cx[k + 2] != CELL_SOLID) + (cym[k] != CELL_SOLID) + (cyp[k] != CELL_SOLID) +
                (czm[k] != CELL_SOLID) + (czp[k] != CELL_SOLID);
        q.diag[k] = (float)d;
        q.nbr[k] = (cx[k] == CELL_FLUID ? 1u : 0u) | (cx[k + 2] == CELL_FLUID ? 2u : 0u) | (cym[k] == CELL_FLUID ? 4u : 0u) |
                   (cyp[k] == CELL_FLUID ? 8u : 0u) | (czm[k] == CELL_FLUID ? 16u : 0u) | (czp[k] == CELL_FLUID ? 32u : 0u);
    }
    return true;
}

// (A x) for the four cells of a quad: MultiplyWithCoefficientMatrix, pressure.glsl:34-75
__device__ __forceinline__ void apply_coeff(const float *__restrict__ x, int64_t i, int sy, int sz, const QuadStencil &q,
                                            const float xc[4], float out[4]) {
    const float4 ym = *reinterpret_cast<const float4 *>(x + i - sy), yp = *reinterpret_cast<const float4 *>(x + i + sy);
    const float4 zm = *reinterpret_cast<const float4 *>(x + i - sz), zp = *reinterpret_cast<const float4 *>(x + i + sz);
    const float xr[6] = {x[i - 1], xc[0], xc[1], xc[2], xc[3], x[i + 4]};
    const float vym[4] = {ym.x, ym.y, ym.z, ym.w}, vyp[4] = {yp.x, yp.y, yp.z, yp.w};
    const float vzm[4] = {zm.x, zm.y, zm.z, zm.w}, vzp[4] = {zp.x, zp.y, zp.z, zp.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float r = q.diag[k] * xc[k];
        const unsigned n = q.nbr[k];
        if (n & 1u) r -= xr[k];
        if (n & 2u) r -= xr[k + 2];
        if (n & 4u) r -= vym[k];
        if (n & 8u) r -= vyp[k];
        if (n & 16u) r -= vzm[k];
        if (n & 32u) r -= vzp[k];
        out[k] = r;
    }
}

// z = P(P(r)) with the LOD-1 neighbour fetches reading 0: z = (r / d) / d, d = max(diag, 1)
// (pressure_apply_preconditioner.comp:48-77).  1/d^2 by table: <= 1.5 ulp from the two divisions.
__device__ __forceinline__ float precond_diag2(float r, float diag) {
    const float inv[7] = {1.0f, 1.0f, 0.25f, 1.0f / 9.0f, 0.0625f, 0.04f, 1.0f / 36.0f};
    return r * inv[(int)diag];
}

__device__ __forceinline__ float block_sum(float v, float *sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = 0.0f;
    if (w == 0) {
        r = lane < (PCG_THREADS / 32) ? sh[lane] : 0.0f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r; // valid in warp 0
}
__device__ __forceinline__ float block_max(float v, float *sh) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    float r = 0.0f;
    if (w == 0) {
        r = lane < (PCG_THREADS / 32) ? sh[lane] : 0.0f;
        r = warp_max(r);
    }
    __syncthreads();
    return r;
}

// Publishes this block's partial(s) and returns true in exactly one block: the last one to arrive.
__device__ __forceinline__ bool publish_and_elect(PcgScalars *scal, unsigned nblocks) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(&scal->ticket, 1u);
        last = (t == nblocks - 1);
        __threadfence(); // acquire side: the partials of all earlier arrivals are visible to this block from here on
    }
    __syncthreads();
    return last;
}

// Final reduction by the elected block, in fixed order and double precision => bit-reproducible run to run.
__device__ __forceinline__ double final_sum(const float *partials, int n, float *sh_unused) {
    __shared__ double shd[PCG_THREADS / 32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < n; k += PCG_THREADS) acc += (double)__ldcg(partials + k);
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) shd[w] = acc;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < (PCG_THREADS / 32) ? shd[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}
__device__ __forceinline__ float final_max(const float *partials, int n, float *sh) {
    float acc = 0.0f;
    for (int k = threadIdx.x; k < n; k += PCG_THREADS) acc = fmaxf(acc, __ldcg(partials + k));
    return block_max(acc, sh);
}

__device__ __forceinline__ float guarded_div(float num, float den) { // pressure_reduce.comp:73-80
    return num / (den + (den < 0.0f ? -PCG_EPSILON : PCG_EPSILON));
}

// ---------------------------------------------------------------------------------------------------------------
// init: p <- 0 off-fluid, r <- b - A p (warm start), and for the diag2 preconditioner s <- z, sigma <- z.r
// (pressure_init.comp:19-84 + the init block of pressure_solver.rs:625-649)
template <int MODE>
__global__ void __launch_bounds__(PCG_THREADS) pcg_init_kernel(GridDim g, TileMap t, const int8_t *__restrict__ marker,
                                                               float *__restrict__ p, float *__restrict__ r,
                                                               float *__restrict__ s, PcgScalars *scal, float *partials) {
    __shared__ float sh[PCG_THREADS / 32];
    float acc = 0.0f;
    int64_t i;
    if (tile_cell(g, t, i)) {
        QuadStencil q;
        if (!load_quad_stencil(marker, i, g.sy, g.sz, q)) {
            *reinterpret_cast<float4 *>(p + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const float4 p4 = *reinterpret_cast<const float4 *>(p + i);
            const float4 b4 = *reinterpret_cast<const float4 *>(r + i);
            float pc[4] = {p4.x, p4.y, p4.z, p4.w}, rc[4] = {b4.x, b4.y, b4.z, b4.w}, Ap[4], sc[4];
            apply_coeff(p, i, g.sy, g.sz, q, pc, Ap);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (q.fluid[k]) {
                    rc[k] -= Ap[k];
                    if (MODE == 0) {
                        float z = precond_diag2(rc[k], q.diag[k]);
                        sc[k] = z;
                        acc += z * rc[k];
                    }
                } else {
                    pc[k] = 0.0f;
                    sc[k] = 0.0f;
                }
            }
            *reinterpret_cast<float4 *>(p + i) = make_float4(pc[0], pc[1], pc[2], pc[3]);
            *reinterpret_cast<float4 *>(r + i) = make_float4(rc[0], rc[1], rc[2], rc[3]);
            if (MODE == 0) *reinterpret_cast<float4 *>(s + i) = make_float4(sc[0], sc[1], sc[2], sc[3]);
        }
    }
    if (MODE != 0) return;
    float bs = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    if (publish_and_elect(scal, t.nblocks)) {
        double tot = final_sum(partials, t.nblocks, sh);
        if (threadIdx.x == 0) {
            scal->alpha = 0.0f; // RESULTMODE_INIT, pressure_reduce.comp:68-71
            scal->beta = 0.0f;
            scal->sigma = (float)tot;
            scal->ticket = 0u;
        }
    }
}

// One half-pass of the preconditioner as literally written with the LOD clamped to 0 (MODE 1 only):
// out = (in - sum over FLUID -x,-y,-z neighbours of in) / diag;  pass 1 also reduces out . r
// (pressure_apply_preconditioner.comp:36-82).  result_mode: 1 = INIT (sigma), 3 = BETA.
__global__ void __launch_bounds__(PCG_THREADS) pcg_precond_pass_kernel(GridDim g, TileMap t, const int8_t *__restrict__ marker,
                                                                       const float *__restrict__ in, float *__restrict__ out,
                                                                       const float *__restrict__ r, int pass1, int result_mode,
                                                                       PcgScalars *scal, float *partials) {
    __shared__ float sh[PCG_THREADS / 32];
    if (scal->done) return;
    float acc = 0.0f;
    int64_t i;
    if (tile_cell(g, t, i)) {
        QuadStencil q;
        if (load_quad_stencil(marker, i, g.sy, g.sz, q)) {
            const float4 c4 = *reinterpret_cast<const float4 *>(in + i);
            const float4 ym = *reinterpret_cast<const float4 *>(in + i - g.sy), zm = *reinterpret_cast<const float4 *>(in + i - g.sz);
            const float xr[5] = {in[i - 1], c4.x, c4.y, c4.z, c4.w};
            const float vym[4] = {ym.x, ym.y, ym.z, ym.w}, vzm[4] = {zm.x, zm.y, zm.z, zm.w};
            float4 o4 = *reinterpret_cast<const float4 *>(out + i);
            float oc[4] = {o4.x, o4.y, o4.z, o4.w};
            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pass1) r4 = *reinterpret_cast<const float4 *>(r + i);
            const float rc[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!q.fluid[k]) continue;
                float v = xr[k + 1];
                if (q.nbr[k] & 1u) v -= xr[k];
                if (q.nbr[k] & 4u) v -= vym[k];
                if (q.nbr[k] & 16u) v -= vzm[k];
                if (q.diag[k] > 0.0f) v /= q.diag[k];
                oc[k] = v;
                acc += v * rc[k];
            }
            *reinterpret_cast<float4 *>(out + i) = make_float4(oc[0], oc[1], oc[2], oc[3]);
        }
    }
    if (!pass1) return;
    float bs = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    if (publish_and_elect(scal, t.nblocks)) {
        double tot = final_sum(partials, t.nblocks, sh);
        if (threadIdx.x == 0) {
            float zr = (float)tot;
            if (result_mode == 1) {
                scal->alpha = 0.0f;
                scal->beta = 0.0f;
                scal->sigma = zr;
            } else {
                scal->beta = guarded_div(zr, scal->sigma);
                scal->sigma = zr;
            }
            scal->ticket = 0u;
        }
    }
}

// dot: alpha <- sigma / (s . A s)   (pressure_apply_coeff.comp:19-30 + RESULTMODE_ALPHA)
__global__ void __launch_bounds__(PCG_THREADS) pcg_dot_kernel(GridDim g, TileMap t, const int8_t *__restrict__ marker,
                                                              const float *__restrict__ s, PcgScalars *scal, float *partials) {
    __shared__ float sh[PCG_THREADS / 32];
    if (scal->done) return;
    float acc = 0.0f;
    int64_t i;
    if (tile_cell(g, t, i)) {
        QuadStencil q;
        if (load_quad_stencil(marker, i, g.sy, g.sz, q)) {
            const float4 s4 = *reinterpret_cast<const float4 *>(s + i);
            const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
            float As[4];
            apply_coeff(s, i, g.sy, g.sz, q, sc, As);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (q.fluid[k]) acc += sc[k] * As[k];
        }
    }
    float bs = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    if (publish_and_elect(scal, t.nblocks)) {
        double tot = final_sum(partials, t.nblocks, sh);
        if (threadIdx.x == 0) {
            scal->alpha = guarded_div(scal->sigma, (float)tot);
            scal->ticket = 0u;
        }
    }
}

// update: p += alpha s; r -= alpha A s; [MODE 0: beta, sigma <- z.r]; [WITH_ERR: max|r| -> statistics / done]
// (pressure_update_pressure_and_residual.comp:23-59, reduce MAX_ERROR pressure_reduce.comp:82-94, and for MODE 0
//  both preconditioner passes + reduce BETA)
template <int MODE, bool WITH_ERR>
__global__ void __launch_bounds__(PCG_THREADS) pcg_update_kernel(GridDim g, TileMap t, const int8_t *__restrict__ marker,
                                                                 float *__restrict__ p, float *__restrict__ r,
                                                                 const float *__restrict__ s, PcgScalars *scal, float *partials,
                                                                 const StepParams *__restrict__ params, int which, int iteration,
                                                                 int max_iterations) {
    __shared__ float sh[PCG_THREADS / 32];
    if (scal->done) return;
    const float alpha = scal->alpha;
    float acc = 0.0f, err = 0.0f;
    int64_t i;
    if (tile_cell(g, t, i)) {
        QuadStencil q;
        if (load_quad_stencil(marker, i, g.sy, g.sz, q)) {
            const float4 s4 = *reinterpret_cast<const float4 *>(s + i);
            const float4 p4 = *reinterpret_cast<const float4 *>(p + i);
            const float4 r4 = *reinterpret_cast<const float4 *>(r + i);
            const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
            float pc[4] = {p4.x, p4.y, p4.z, p4.w}, rc[4] = {r4.x, r4.y, r4.z, r4.w}, As[4];
            apply_coeff(s, i, g.sy, g.sz, q, sc, As);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!q.fluid[k]) continue;
                pc[k] = pc[k] + alpha * sc[k];
                rc[k] -= alpha * As[k];
                if (MODE == 0) acc += precond_diag2(rc[k], q.diag[k]) * rc[k];
                if (WITH_ERR) err = fmaxf(err, fabsf(rc[k]));
            }
            *reinterpret_cast<float4 *>(p + i) = make_float4(pc[0], pc[1], pc[2], pc[3]);
            *reinterpret_cast<float4 *>(r + i) = make_float4(rc[0], rc[1], rc[2], rc[3]);
        }
    }
    if (MODE != 0 && !WITH_ERR) return;
    float *pmax = partials + t.nblocks;
    if (MODE == 0) {
        float bs = block_sum(acc, sh);
        if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    }
    if (WITH_ERR) {
        float bm = block_max(err, sh);
        if (threadIdx.x == 0) pmax[blockIdx.x] = bm;
    }
    if (publish_and_elect(scal, t.nblocks)) {
        double tot = 0.0;
        float e = 0.0f;
        if (MODE == 0) tot = final_sum(partials, t.nblocks, sh);
        if (WITH_ERR) e = final_max(pmax, t.nblocks, sh);
        if (threadIdx.x == 0) {
            if (WITH_ERR) {
                const float tol = params->tolerance[which];
                if (scal->num_iterations == 0 && (iteration == max_iterations || e < tol)) {
                    scal->max_error = e;
                    scal->num_iterations = iteration;
                    scal->done = 1;
                }
            }
            if (MODE == 0) {
                float zr = (float)tot;
                scal->beta = guarded_div(zr, scal->sigma);
                scal->sigma = zr;
            }
            scal->ticket = 0u;
        }
    }
}

// search: s <- z + beta s on fluid cells (pressure_update_search.comp:13-24); MODE 0 recomputes z = r / diag^2
template <int MODE>
__global__ void __launch_bounds__(PCG_THREADS) pcg_search_kernel(GridDim g, TileMap t, const int8_t *__restrict__ marker,
                                                                 float *__restrict__ s, const float *__restrict__ r_or_z,
                                                                 const PcgScalars *scal) {
    if (scal->done) return;
    const float beta = scal->beta;
    int64_t i;
    if (!tile_cell(g, t, i)) return;
    if (MODE == 0) {
        QuadStencil q;
        if (!load_quad_stencil(marker, i, g.sy, g.sz, q)) return;
        const float4 s4 = *reinterpret_cast<const float4 *>(s + i);
        const float4 r4 = *reinterpret_cast<const float4 *>(r_or_z + i);
        float sc[4] = {s4.x, s4.y, s4.z, s4.w};
        const float rc[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (q.fluid[k]) sc[k] = precond_diag2(rc[k], q.diag[k]) + beta * sc[k];
        *reinterpret_cast<float4 *>(s + i) = make_float4(sc[0], sc[1], sc[2], sc[3]);
    } else {
        const char4 c = *reinterpret_cast<const char4 *>(marker + i);
        if (c.x != CELL_FLUID && c.y != CELL_FLUID && c.z != CELL_FLUID && c.w != CELL_FLUID) return;
        const float4 s4 = *reinterpret_cast<const float4 *>(s + i);
        const float4 z4 = *reinterpret_cast<const float4 *>(r_or_z + i);
        float4 o = s4;
        if (c.x == CELL_FLUID) o.x = z4.x + beta * s4.x;
        if (c.y == CELL_FLUID) o.y = z4.y + beta * s4.y;
        if (c.z == CELL_FLUID) o.z = z4.z + beta * s4.z;
        if (c.w == CELL_FLUID) o.w = z4.w + beta * s4.w;
        *reinterpret_cast<float4 *>(s + i) = o;
    }
}

__global__ void pcg_reset_scalars_kernel(PcgScalars *scal) {
    scal->alpha = 0.0f; scal->beta = 0.0f; scal->sigma = 0.0f;
    scal->max_error = 0.0f; scal->num_iterations = 0; scal->done = 0; scal->ticket = 0u;
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------
PressureField::PressureField(const GridDim &grid, const SolverConfig &cfg) : config(cfg) {
    pressure_.alloc(grid);
    BLUB_CUDA_CHECK(cudaMalloc(&scalars, sizeof(PcgScalars)));
    BLUB_CUDA_CHECK(cudaMemset(scalars, 0, sizeof(PcgScalars)));
    BLUB_CUDA_CHECK(cudaMallocHost(&pinned_, sizeof(float) * 2 * NUM_PRESSURE_ERROR_BUFFER));
    ring_.resize(NUM_PRESSURE_ERROR_BUFFER);
    for (int k = 0; k < NUM_PRESSURE_ERROR_BUFFER; ++k) {
        BLUB_CUDA_CHECK(cudaEventCreateWithFlags(&ring_[k].event, cudaEventDisableTiming));
        ring_[k].host = pinned_ + 2 * k;
        ring_[k].in_flight = false;
        unused_.push_back(k);
    }
}

PressureField::~PressureField() {
    for (auto &p : ring_) cudaEventDestroy(p.event);
    if (pinned_) cudaFreeHost(pinned_);
    if (scalars) cudaFree(scalars);
    pressure_.release();
}

void PressureField::retrieve_new_error_samples() {
    while (!pending_.empty()) {
        Pending &pb = ring_[pending_.front()];
        if (cudaEventQuery(pb.event) != cudaSuccess) break; // oldest first; later ones cannot be done either
        SolverStatisticSample smp;
        // "We always deal with pressure * dt / density": scale the error by dt for display (pressure_solver.rs:158-163)
        smp.error = pb.host[0] * pb.dt;
        smp.iteration_count = (int32_t)pb.host[1];
        stats.push_back(smp);
        while (stats.size() > SOLVER_STATISTIC_HISTORY_LENGTH) stats.pop_front();
        pb.in_flight = false;
        unused_.push_back(pending_.front());
        pending_.pop_front();
    }
}

namespace {
__global__ void pcg_export_stats_kernel(const PcgScalars *scal, float *out) {
    out[0] = scal->max_error;
    out[1] = (float)scal->num_iterations;
}
} // namespace

void PressureField::enqueue_error_buffer_read(cudaStream_t stream, float simulation_delta) {
    if (unused_.empty()) return; // "No more error buffer available for async copy" (pressure_solver.rs:188-190)
    int k = unused_.back();
    unused_.pop_back();
    Pending &pb = ring_[k];
    // pinned host memory is device-accessible under UVA: the 8-byte result is stored straight into it
    BLUB_LAUNCH(pcg_export_stats_kernel, 1, 1, 0, stream, scalars, pb.host);
    BLUB_CUDA_CHECK(cudaEventRecord(pb.event, stream));
    pb.dt = simulation_delta;
    pb.in_flight = true;
    pending_.push_back(k);
}

void PressureField::read_last_solve(cudaStream_t stream, float *max_error, int *iterations) {
    PcgScalars h;
    BLUB_CUDA_CHECK(cudaMemcpyAsync(&h, scalars, sizeof(h), cudaMemcpyDeviceToHost, stream));
    BLUB_CUDA_CHECK(cudaStreamSynchronize(stream));
    *max_error = h.max_error;
    *iterations = h.num_iterations;
}

PressureSolver::PressureSolver(const GridDim &grid) : grid_(grid) {
    residual_.alloc(grid);
    search_.alloc(grid);
    aux_.alloc(grid);
    aux_temp_.alloc(grid);
    TileMap t = make_tilemap(grid);
    num_blocks_ = t.nblocks;
    BLUB_CUDA_CHECK(cudaMalloc(&partials_, sizeof(float) * 2 * (size_t)num_blocks_));
}

PressureSolver::~PressureSolver() {
    residual_.release(); search_.release(); aux_.release(); aux_temp_.release();
    if (partials_) cudaFree(partials_);
}

void PressureSolver::solve(cudaStream_t stream, PressureField &field, int which, const int8_t *marker, const StepParams *dparams,
                           const Quirks &quirks) {
    const GridDim g = grid_;
    const TileMap t = make_tilemap(g);
    float *p = field.pressure(), *r = residual_.ptr, *s = search_.ptr;
    PcgScalars *scal = field.scalars;
    const int mode = quirks.precond_mode;
    const int max_it = field.config.max_num_iterations;
    const int freq = field.config.error_check_frequency > 0 ? field.config.error_check_frequency : 1;

    field.retrieve_new_error_samples(); // pressure_solver.rs:614
    field.touched = true;               // the volume is zero-initialised at allocation (:601-603)

    BLUB_LAUNCH(pcg_reset_scalars_kernel, 1, 1, 0, stream, scal);
    if (mode == 0) {
        BLUB_LAUNCH(pcg_init_kernel<0>, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_);
    } else {
        BLUB_LAUNCH(pcg_init_kernel<1>, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_);
        BLUB_LAUNCH(pcg_precond_pass_kernel, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, r, aux_temp_.ptr, r, 0, 0, scal, partials_);
        BLUB_LAUNCH(pcg_precond_pass_kernel, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, aux_temp_.ptr, s, r, 1, 1, scal, partials_);
    }
    for (int i = 0;; ++i) {
        BLUB_LAUNCH(pcg_dot_kernel, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, s, scal, partials_);
        const bool with_err = (max_it == i) || (i > 0 && i % freq == 0); // pressure_solver.rs:676-677
        if (mode == 0) {
            if (with_err)
                BLUB_LAUNCH((pcg_update_kernel<0, true>), t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_, dparams, which, i, max_it);
            else
                BLUB_LAUNCH((pcg_update_kernel<0, false>), t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_, dparams, which, i, max_it);
        } else {
            if (with_err)
                BLUB_LAUNCH((pcg_update_kernel<1, true>), t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_, dparams, which, i, max_it);
            else
                BLUB_LAUNCH((pcg_update_kernel<1, false>), t.nblocks, PCG_THREADS, 0, stream, g, t, marker, p, r, s, scal, partials_, dparams, which, i, max_it);
        }
        if (i >= max_it) break; // :699-701
        if (mode == 0) {
            BLUB_LAUNCH(pcg_search_kernel<0>, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, s, r, scal);
        } else {
            BLUB_LAUNCH(pcg_precond_pass_kernel, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, r, aux_temp_.ptr, r, 0, 0, scal, partials_);
            BLUB_LAUNCH(pcg_precond_pass_kernel, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, aux_temp_.ptr, aux_.ptr, r, 1, 3, scal, partials_);
            BLUB_LAUNCH(pcg_search_kernel<1>, t.nblocks, PCG_THREADS, 0, stream, g, t, marker, s, aux_.ptr, scal);
        }
    }
}

} // namespace blub
