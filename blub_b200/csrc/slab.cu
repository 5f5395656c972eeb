// slab.cu -- z-slab sharding of the WHOLE step across GPUs (the reference is single-GPU).
//
// Rank k owns the global planes [k*zs, (k+1)*zs); its local grid carries SLAB_HALO ghost planes on both sides and all
// stage kernels run unchanged on that local grid.  What crosses the slab faces, per step:
//   X1  after the P2G scatter      SUM of the (num, weight) accumulators and MAX of the markers over the 4 planes around
//                                  each face (both ranks send their partials of the same 4 global planes and add)
//   PCG                            boundary planes of r / p pushed from inside the persistent kernel (pcg.cu)
//   X2  after extrapolation        COPY of the first two owned planes of u into the neighbour's ghost planes
//   MIG after advection            particles that left the slab move to the neighbour (positions re-based by +-zs)
//   X3  after advection            MAX of the markers
//   X4  after the density scatter  SUM of the density accumulator
//   X5  after extrapolation #2     COPY of the displacement field planes
// Transport: stream-ordered cudaMemcpyAsync straight into the neighbour's (double-buffered) receive area inside its
// peer-visible window, one mailbox barrier kernel per exchange, then a local combine kernel.  No host synchronisation,
// no NCCL.  Particles keep their owner until they cross a face; up to half a cell of overhang (density correction) is
// tolerated by the 4-plane overlap.
#include <cstring>

#include "blub_core.hpp"
#include "fluid_kernels.hpp"

namespace blub {
namespace {

constexpr long long BARRIER_SPIN_LIMIT = 40LL * 1000 * 1000;

// one mailbox round (same protocol and the same round counter as comm_allreduce in pcg.cu): stream-ordered barrier
__global__ void slab_barrier_kernel(SlabComm c, int *error_flag) {
    __shared__ int dead;
    const int tid = threadIdx.x;
    if (tid == 0) dead = 0;
    __syncthreads();
    const unsigned seq = *c.seq + 1u;
    const int slot = (int)(seq & 1u) * 2 * SLAB_MAX_WORLD;
    if (tid < c.world) {
        __threadfence_system();
        volatile unsigned long long *dst = c.mailbox[tid] + slot + 2 * c.rank;
        dst[0] = (unsigned long long)seq << 32;
        dst[1] = (unsigned long long)seq << 32;
        volatile unsigned long long *src = c.mailbox[c.rank] + slot + 2 * tid;
        long long spins = 0;
        while ((unsigned)(src[0] >> 32) != seq || (unsigned)(src[1] >> 32) != seq) {
            if (++spins > BARRIER_SPIN_LIMIT) { dead = 1; break; }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid == 0) {
        *c.seq = seq;
        if (dead && error_flag) *error_flag = 1;
    }
}

__global__ void __launch_bounds__(256) halo_add_kernel(float *__restrict__ dst, const float *__restrict__ src, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
__global__ void __launch_bounds__(256) halo_max_kernel(int8_t *__restrict__ dst, const int8_t *__restrict__ src, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) dst[i] = max(dst[i], src[i]);
}

struct MigrantRecord { // what travels: 64 B per particle
    float4 pos, rx, ry, rz;
};

// counters: [0] stay, [1] down, [2] up, [3] overflow flag
__global__ void __launch_bounds__(256) migrate_classify_kernel(const StepParams *__restrict__ params, const float4 *__restrict__ pos,
                                                               const float4 *__restrict__ rx, const float4 *__restrict__ ry,
                                                               const float4 *__restrict__ rz, float4 *__restrict__ opos, float4 *__restrict__ orx,
                                                               float4 *__restrict__ ory, float4 *__restrict__ orz, MigrantRecord *peer_down,
                                                               MigrantRecord *peer_up, unsigned int *counters, float z_lo, float z_hi, float zshift,
                                                               unsigned int capacity) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= params->num_particles) return;
    float4 p = pos[i];
    const int dest = (p.z < z_lo && peer_down) ? 1 : ((p.z >= z_hi && peer_up) ? 2 : 0);
    // warp-aggregated slot claim: one atomic per destination per warp
    const unsigned active = __activemask();
    unsigned slot = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const unsigned m = __ballot_sync(active, dest == d);
        if (m == 0u) continue;
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counters + d, (unsigned)__popc(m));
        base = __shfl_sync(active, base, leader);
        if (dest == d) slot = base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
    }
    if (dest == 0) {
        opos[slot] = p;
        orx[slot] = rx[i];
        ory[slot] = ry[i];
        orz[slot] = rz[i];
    } else {
        if (slot >= capacity) { // cannot happen with a CFL-limited flow; never overrun the neighbour's buffer
            counters[3] = 1u;
            return;
        }
        p.z += dest == 1 ? zshift : -zshift; // re-base into the neighbour's local frame
        MigrantRecord rec = {p, rx[i], ry[i], rz[i]};
        (dest == 1 ? peer_down : peer_up)[slot] = rec; // P2P store over NVLink
    }
}

// publish how many particles were sent (into the neighbours' windows) and start the new local count
__global__ void migrate_publish_kernel(unsigned int *counters, unsigned int *peer_count_down, unsigned int *peer_count_up, unsigned int capacity) {
    if (threadIdx.x != 0) return;
    if (peer_count_down) *peer_count_down = min(counters[1], capacity);
    if (peer_count_up) *peer_count_up = min(counters[2], capacity);
    __threadfence_system();
}

__global__ void __launch_bounds__(256) migrate_append_kernel(const MigrantRecord *__restrict__ recv_lo, const MigrantRecord *__restrict__ recv_hi,
                                                             const unsigned int *__restrict__ recv_counts, const unsigned int *__restrict__ counters,
                                                             float4 *__restrict__ opos, float4 *__restrict__ orx, float4 *__restrict__ ory,
                                                             float4 *__restrict__ orz, unsigned int max_particles) {
    const unsigned n_lo = recv_counts[0], n_hi = recv_counts[1], stay = counters[0];
    const unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_lo + n_hi) return;
    const MigrantRecord rec = i < n_lo ? recv_lo[i] : recv_hi[i - n_lo];
    const unsigned dst = stay + i;
    if (dst >= max_particles) return;
    opos[dst] = rec.pos;
    orx[dst] = rec.rx;
    ory[dst] = rec.ry;
    orz[dst] = rec.rz;
}

__global__ void migrate_finish_kernel(StepParams *params, unsigned int *counters, unsigned int *recv_counts, unsigned int max_particles, int *error_flag) {
    if (threadIdx.x != 0) return;
    unsigned n = counters[0] + recv_counts[0] + recv_counts[1];
    if (n > max_particles) {
        n = max_particles;
        if (error_flag) *error_flag = 2;
    }
    if (counters[3] && error_flag) *error_flag = 3;
    params->num_particles = n;
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    recv_counts[0] = recv_counts[1] = 0;
}

inline int blocks_for(int64_t n, int per) { return (int)((n + per - 1) / per); }

} // namespace

// ---------------------------------------------------------------------------------------------------------------
size_t HybridFluid::slab_extra_window_bytes() const {
    const size_t cols = (size_t)grid_.nx * grid_.ny;
    const size_t halo = (cols * 100 + 255) / 256 * 256;                              // 3 x float2 x 4 planes + marker x 4 planes
    const size_t part = ((size_t)slab_migrant_capacity() * sizeof(MigrantRecord) + 255) / 256 * 256;
    return 4 * halo + 2 * part + 256;
}

uint32_t HybridFluid::slab_migrant_capacity() const { return (uint32_t)grid_.nx * grid_.ny * 8u * 3u; } // three planes' worth

void HybridFluid::slab_layout(void *window, char *&halo0, size_t &halo_bytes, char *&part0, size_t &part_bytes, unsigned int *&counts) const {
    const size_t vol = GridArray<float>::bytes_for(grid_);
    const size_t cols = (size_t)grid_.nx * grid_.ny;
    halo_bytes = (cols * 100 + 255) / 256 * 256;
    part_bytes = ((size_t)slab_migrant_capacity() * sizeof(MigrantRecord) + 255) / 256 * 256;
    char *w = static_cast<char *>(window) + 4096 + 3 * vol;
    halo0 = w;
    part0 = w + 4 * halo_bytes;
    counts = reinterpret_cast<unsigned int *>(part0 + 2 * part_bytes);
}

void HybridFluid::slab_barrier() {
    BLUB_LAUNCH(slab_barrier_kernel, 1, 32, 0, stream_, solver_->comm, slab_error_);
}

// items: {cell-0 pointer, bytes per cell}.  kind 0: SUM over the 4 overlap planes (float data), 1: MAX (int8), 2: COPY of the
// two outermost owned planes into the neighbour's two innermost ghost planes.
void HybridFluid::slab_halo_exchange(const SlabHaloItem *items, int n_items) {
    const SlabComm &c = solver_->comm;
    if (c.world <= 1) return;
    const int H = SLAB_HALO, zs = c.owned_nz;
    const size_t plane_cells = (size_t)grid_.nx * grid_.ny;
    const int buf = (int)(slab_exchange_index_++ & 1u);
    char *halo0;
    size_t hb, pb;
    char *part0;
    unsigned int *counts;
    slab_layout(window_, halo0, hb, part0, pb, counts);
    // receive areas: [face 0 = from the lower neighbour, face 1 = from the upper neighbour][buffer]
    auto local_recv = [&](int face) { return halo0 + (size_t)(face * 2 + buf) * hb; };
    auto peer_recv = [&](int side) -> char * { // my message to the neighbour on `side` lands in ITS area for the opposite face
        if (!slab_peer_window_[side]) return nullptr;
        char *ph0, *pp0;
        size_t phb, ppb;
        unsigned int *pc;
        slab_layout(slab_peer_window_[side], ph0, phb, pp0, ppb, pc);
        return ph0 + (size_t)((1 - side) * 2 + buf) * phb;
    };
    // 1) send
    for (int side = 0; side < 2; ++side) {
        char *dst = peer_recv(side);
        if (!dst) continue;
        size_t off = 0;
        for (int k = 0; k < n_items; ++k) {
            const SlabHaloItem &it = items[k];
            const int planes = it.kind == 2 ? 2 : 4;
            int first;
            if (it.kind == 2) first = side == 0 ? H : H + zs - 2;        // my outermost owned planes
            else first = side == 0 ? H - 2 : H + zs - 2;                 // the 4 planes around the face
            const size_t bytes = (size_t)planes * plane_cells * it.bytes_per_cell;
            const char *src = static_cast<const char *>(it.cell0) + (size_t)first * plane_cells * it.bytes_per_cell;
            BLUB_CUDA_CHECK(cudaMemcpyAsync(dst + off, src, bytes, cudaMemcpyDeviceToDevice, stream_));
            off += (bytes + 255) / 256 * 256;
        }
    }
    // 2) everybody's messages have landed
    slab_barrier();
    // 3) combine
    for (int face = 0; face < 2; ++face) {
        const int nb = slab_rank_ + (face == 0 ? -1 : 1);
        if (nb < 0 || nb >= c.world) continue;
        const char *src = local_recv(face);
        size_t off = 0;
        for (int k = 0; k < n_items; ++k) {
            const SlabHaloItem &it = items[k];
            const int planes = it.kind == 2 ? 2 : 4;
            int first;
            if (it.kind == 2) first = face == 0 ? H - 2 : H + zs;        // my innermost ghost planes
            else first = face == 0 ? H - 2 : H + zs - 2;
            const size_t bytes = (size_t)planes * plane_cells * it.bytes_per_cell;
            char *dst = static_cast<char *>(it.cell0) + (size_t)first * plane_cells * it.bytes_per_cell;
            if (it.kind == 0) {
                const int64_t n = (int64_t)(bytes / sizeof(float));
                BLUB_LAUNCH(halo_add_kernel, blocks_for(n, 256), 256, 0, stream_, reinterpret_cast<float *>(dst), reinterpret_cast<const float *>(src + off), n);
            } else if (it.kind == 1) {
                BLUB_LAUNCH(halo_max_kernel, blocks_for((int64_t)bytes, 256), 256, 0, stream_, reinterpret_cast<int8_t *>(dst),
                            reinterpret_cast<const int8_t *>(src + off), (int64_t)bytes);
            } else {
                BLUB_CUDA_CHECK(cudaMemcpyAsync(dst, src + off, bytes, cudaMemcpyDeviceToDevice, stream_));
            }
            off += (bytes + 255) / 256 * 256;
        }
    }
}

// particles that crossed a slab face move to the neighbour; the survivors are compacted into the spare arrays
void HybridFluid::slab_migrate() {
    const SlabComm &c = solver_->comm;
    if (c.world <= 1) return;
    char *halo0, *part0;
    size_t hb, pb;
    unsigned int *counts;
    slab_layout(window_, halo0, hb, part0, pb, counts);
    MigrantRecord *peer_buf[2] = {nullptr, nullptr};
    unsigned int *peer_cnt[2] = {nullptr, nullptr};
    for (int side = 0; side < 2; ++side) {
        if (!slab_peer_window_[side]) continue;
        char *ph0, *pp0;
        size_t phb, ppb;
        unsigned int *pc;
        slab_layout(slab_peer_window_[side], ph0, phb, pp0, ppb, pc);
        peer_buf[side] = reinterpret_cast<MigrantRecord *>(pp0 + (size_t)(1 - side) * ppb); // lands in its area for the opposite face
        peer_cnt[side] = pc + (1 - side);
    }
    const uint32_t cap = slab_migrant_capacity();
    const uint32_t np_upper = max_num_particles_;
    float4 *opos = pos_[1 - cur_];
    BLUB_LAUNCH(migrate_classify_kernel, blocks_for(np_upper, 256), 256, 0, stream_, params_dev_, pos_[cur_], row_[0], row_[1], row_[2], opos,
                row_alt_[0], row_alt_[1], row_alt_[2], peer_buf[0], peer_buf[1], mig_counters_, (float)SLAB_HALO, (float)(SLAB_HALO + c.owned_nz),
                (float)c.owned_nz, cap);
    BLUB_LAUNCH(migrate_publish_kernel, 1, 32, 0, stream_, mig_counters_, peer_cnt[0], peer_cnt[1], cap);
    slab_barrier();
    BLUB_LAUNCH(migrate_append_kernel, blocks_for(2 * (int64_t)cap, 256), 256, 0, stream_, reinterpret_cast<const MigrantRecord *>(part0),
                reinterpret_cast<const MigrantRecord *>(part0 + pb), counts, mig_counters_, opos, row_alt_[0], row_alt_[1], row_alt_[2], max_num_particles_);
    BLUB_LAUNCH(migrate_finish_kernel, 1, 32, 0, stream_, params_dev_, mig_counters_, counts, max_num_particles_, slab_error_);
    cur_ = 1 - cur_;
    for (int k = 0; k < 3; ++k) std::swap(row_[k], row_alt_[k]);
    row_parity_ ^= 1;
}

} // namespace blub
