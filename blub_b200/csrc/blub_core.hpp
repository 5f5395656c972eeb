// blub_core.hpp -- C++ host side of libblubcore: the B200-native counterparts of the reference's
// `HybridFluid` (src/simulation/hybrid_fluid.rs), `PressureSolver` and `PressureField`
// (src/simulation/pressure_solver.rs).  Same method names, argument meaning and defaults; wgpu objects are replaced by
// device pointers and one CUDA stream.  The C ABI in include/blub_fluid.h is a thin shim over these classes.
#pragma once
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/blub_fluid.h"
#include "common.cuh"
#include "fluid_kernels.hpp"

namespace blub {

// pressure_solver.rs:57-62
struct SolverConfig {
    float error_tolerance = 0.1f;
    int32_t max_num_iterations = 32;
    int32_t error_check_frequency = 4;
};
// pressure_solver.rs:63-68
struct SolverStatisticSample {
    float error = 0.0f;
    int32_t iteration_count = 0;
};
// hybrid_fluid.rs:19-22
struct DynamicSettings {
    uint32_t particle_rebinning_step_frequency = 60; // hybrid_fluid.rs:603-605
};

struct Quirks {
    int precond_mode = 0; // 0: z = r/diag^2 ("LOD-1 fetch returns 0"), 1: as written with clamped LOD
};

// A padded device array of grid cells (see GridDim).
template <class T> struct GridArray {
    T *base = nullptr; // allocation
    T *ptr = nullptr;  // cell 0
    bool owned = true;
    static size_t bytes_for(const GridDim &g) { return (((size_t)(g.n + 2 * g.pad) * sizeof(T)) + 255) / 256 * 256; }
    void alloc(const GridDim &g) {
        size_t bytes = (size_t)(g.n + 2 * g.pad) * sizeof(T);
        BLUB_CUDA_CHECK(cudaMalloc(&base, bytes));
        BLUB_CUDA_CHECK(cudaMemset(base, 0, bytes));
        ptr = base + g.pad;
        owned = true;
    }
    // place the array in caller-owned, zero-initialised memory (the peer-visible slab window)
    void place(const GridDim &g, void *mem) {
        base = static_cast<T *>(mem);
        ptr = base + g.pad;
        owned = false;
    }
    void release() {
        if (base && owned) cudaFree(base);
        base = ptr = nullptr;
    }
};

class PressureSolver;

// PressureField, pressure_solver.rs:84-210: one pressure volume (kept between steps = warm start), its solver
// configuration, and the asynchronous {max error, iteration count} read-back ring.
class PressureField {
  public:
    static constexpr size_t SOLVER_STATISTIC_HISTORY_LENGTH = 100; // pressure_solver.rs:101
    static constexpr int NUM_PRESSURE_ERROR_BUFFER = 32;           // pressure_solver.rs:49

    PressureField(const GridDim &grid, const SolverConfig &config, void *external_volume = nullptr);
    ~PressureField();
    PressureField(const PressureField &) = delete;

    SolverConfig config;
    std::deque<SolverStatisticSample> stats;
    float *pressure() const { return pressure_.ptr; }

    // PressureField::retrieve_new_error_samples, pressure_solver.rs:148-174 (never blocks)
    void retrieve_new_error_samples();
    // PressureField::enqueue_error_buffer_read, :176-191
    void enqueue_error_buffer_read(cudaStream_t stream, float simulation_delta);
    // blocking read of the device scalars (tests only)
    void read_last_solve(cudaStream_t stream, float *max_error, int *iterations);

    PcgScalars *scalars = nullptr; // device
    bool touched = false;          // "timestamp_last_iteration == 0" of :601-603

  private:
    struct Pending {
        cudaEvent_t event;
        float *host; // pinned {max_error, float(num_iterations)}
        float dt;
        bool in_flight;
    };
    GridArray<float> pressure_;
    std::vector<Pending> ring_;
    std::deque<int> pending_;
    std::vector<int> unused_;
    float *pinned_ = nullptr;
};

// PressureSolver, pressure_solver.rs:22-47,228-729: scratch volumes shared by both solves and the PCG recording.
class PressureSolver {
  public:
    PressureSolver(const GridDim &grid, void *external_residual = nullptr);
    ~PressureSolver();
    PressureSolver(const PressureSolver &) = delete;

    float *residual() const { return residual_.ptr; } // the rhs is written straight into it (hybrid_fluid.rs:836-838)

    // PressureSolver::solve, pressure_solver.rs:591-729.  Enqueue-only.
    void solve(cudaStream_t stream, PressureField &field, int which, const int8_t *marker, const StepParams *dparams,
               const Quirks &quirks);
    SlabComm comm; // filled by HybridFluid::attach_slab_peers; world == 1 when not sharded
    // {tiles walked by the tile body, quad columns walked by the column body, cells per tile, cells per column} of the last solve (blocking)
    void read_work(cudaStream_t stream, uint32_t out[4]);

  private:
    GridDim grid_;
    GridArray<float> residual_, search_, aux_, aux_temp_;
    GridArray<uint8_t> codes_;       // per-cell code (diag | fluid << 3), rebuilt from the markers at the start of every solve
    uint8_t *tile_active_ = nullptr; // per-tile "contains FLUID" flag
    float *partials_ = nullptr;      // per-tile partial sums / maxima, 2 * tiles (+ slack for the persistent solver)
    int *tile_list_ = nullptr;       // compacted ids of the active tiles
    int *num_active_ = nullptr;
    int num_blocks_ = 0;
    int persistent_blocks_ = 0;      // co-resident blocks of the cooperative solver; 0 = not available

  public:
    bool use_persistent = true;      // one cooperative launch per solve (diag2 preconditioner); false = three kernels per iteration
    bool use_tma = false;            // persistent solver with TMA-staged tiles (nx % 128 == 0); BLUB_PCG=tma or blub_fluid_set_solver_path(f, 2)
    bool use_dense = false;          // persistent solver without the per-thread sparsity skip (comparison only); set_solver_path(f, 4)
    bool use_columns = true;         // single GPU: dense tiles by the tile body, the rest as one list of quad columns (default); BLUB_PCG=tiles
                                     // or set_solver_path(f, 6) walks every active tile with the tile bodies instead
    bool tma_available() const { return tma_blocks_ > 0; }
    bool columns_available() const { return column_blocks_ > 0; }
    bool persistent_available() const { return persistent_blocks_ > 0; }

  private:
    void *tma_maps_ = nullptr;       // PcgTmaMaps (tensor maps of r, s0, s1, codes)
    int tma_blocks_ = 0, column_blocks_ = 0;
    int *tile_cols_ = nullptr;        // column solver: per-tile column counts | exclusive offsets | total
    int *col_list_ = nullptr;         // compacted quad columns (linear index of the quad in its tile's first plane)
    unsigned *barrier_ = nullptr;     // arrival counter of the column solver's grid-wide reductions

  public:
};

// HybridFluid, hybrid_fluid.rs:24-72,92-977
class HybridFluid {
  public:
    static constexpr uint32_t PARTICLES_PER_GRID_CELL = 8; // hybrid_fluid.rs:90

    // slab_world > 1: this fluid is rank `slab_rank` of a z-slab decomposition; `nz` is the number of OWNED planes and the
    // local grid gets SLAB_HALO ghost planes on both sides (see SlabComm).
    HybridFluid(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t max_num_particles, int device, cudaStream_t stream, int slab_rank = 0,
                int slab_world = 1);
    ~HybridFluid();
    HybridFluid(const HybridFluid &) = delete;

    // returns true when the cube was truncated to max_num_particles (hybrid_fluid.rs:627-633)
    bool add_fluid_cube(const float min_grid[3], const float max_grid[3]);
    void set_gravity_grid(const float g[3]) { gravity_[0] = g[0]; gravity_[1] = g[1]; gravity_[2] = g[2]; }
    uint32_t num_particles() const;
    uint32_t num_active_particles() const { return num_particles(); }
    const GridDim &grid_dimension() const { return grid_; }
    SolverConfig &pressure_solver_config_velocity() { return field_velocity_->config; }
    SolverConfig &pressure_solver_config_density() { return field_density_->config; }
    DynamicSettings &dynamic_settings() { return dynamic_settings_; }
    const std::deque<SolverStatisticSample> &pressure_solver_stats_velocity() const { return field_velocity_->stats; }
    const std::deque<SolverStatisticSample> &pressure_solver_stats_density() const { return field_density_->stats; }
    void update_statistics();
    void set_solid_voxels(const void *rgba16f) { voxels_ = static_cast<const uint2 *>(rgba16f); }
    void step(double simulation_delta_seconds);
    void step_stages(double simulation_delta_seconds, int from, int to);
    // one eager step with CUDA events between the 14 stages; ms[14] (synchronises; diagnostics only)
    void step_timed(double simulation_delta_seconds, float ms[14]);
    bool use_graph = true; // replay the step as a CUDA graph (BLUB_NO_GRAPH=1 disables)
    void solve_only(int which, double simulation_delta_seconds);
    void synchronize();
    // peer-visible window (residual, both pressure volumes, mailbox) of a slab rank, and the peers' windows as mapped here
    void *slab_window() const { return window_; }
    size_t slab_window_bytes() const { return window_bytes_; }
    void attach_slab_peers(void *const *windows, int world);
    int slab_rank() const { return slab_rank_; }
    int slab_world() const { return slab_world_; }
    bool sharded() const { return slab_world_ > 1; }
    // 0 = fine; 1 = a peer timed out in a barrier; 2 = particle capacity exceeded; 3 = migration buffer overflow (synchronises)
    int slab_error();
    void invalidate_graphs() { destroy_graphs(); }
    void set_transfer_path(int scatter);
    void marker_written_externally() { fluid_bits_stale_ = true; }

    Quirks quirks;

    // raw views (renderer bind group, hybrid_fluid.rs:351-369) and test taps
    float4 *particles_position() const { return pos_[cur_]; }
    float4 *particles_row(int c) const { return row_[c]; }
    float *grid_velocity(int c) const { return u_[c].ptr; }
    int8_t *marker() const { return marker_.ptr; }
    PressureField &field(int which) { return which == 0 ? *field_velocity_ : *field_density_; }
    PressureSolver &solver() { return *solver_; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    uint32_t max_num_particles() const { return max_num_particles_; }
    void set_particles(uint32_t count, const float *pos4, const float *vx4, const float *vy4, const float *vz4);

  private:
    void upload_step_params(float dt);
    void run_stage(int stage, float dt);
    void refresh_fluid_bits();
    bool binning_step() const;
    uint32_t internal_resort_every_ = 8; // see binning_step(); single GPU and z-slab ranks alike
    void destroy_graphs();

    // ---- z-slab sharding of the whole step (slab.cu) ----
    struct SlabHaloItem {
        void *cell0;           // cell-0 pointer of a grid volume
        size_t bytes_per_cell;
        int kind;              // 0 SUM (float data) / 1 MAX (int8) over the 4 overlap planes, 2 COPY of 2 owned planes into the ghost planes
    };
    size_t slab_extra_window_bytes() const;
    uint32_t slab_migrant_capacity() const;
    void slab_layout(void *window, char *&halo0, size_t &halo_bytes, char *&part0, size_t &part_bytes, unsigned int *&counts) const;
    void slab_barrier();
    void slab_halo_exchange(const SlabHaloItem *items, int n_items);
    MigrateOut slab_migrate_targets();
    void slab_migrate_finish();
    void set_device_particle_count(uint32_t n);
    bool add_fluid_cube_slab(const uint32_t mn[3], const uint32_t ext[3]);
    uint32_t seeded_global_ = 0;
    float4 *row_alt_[3] = {nullptr, nullptr, nullptr}; // spare velocity rows: migration compacts out of place
    unsigned int *mig_counters_ = nullptr;
    int *slab_error_ = nullptr;      // device alias of slab_error_host_
    int *slab_error_host_ = nullptr; // mapped pinned flag: 0 fine, 1 peer timed out, 2 particle capacity exceeded, 3 migration buffer overflow
    void *slab_peer_window_[2] = {nullptr, nullptr};
    uint32_t slab_exchange_index_ = 0;

    // Cached executable graphs of one step, keyed by (position buffer in use, binning step?).  Everything a graph bakes
    // in besides that key is in `GraphSignature`; a change drops the cache.
    struct GraphSignature {
        uint32_t num_particles = 0xFFFFFFFFu;
        const void *voxels = nullptr;
        int precond_mode = -1;
        int max_it[2] = {-1, -1}, freq[2] = {-1, -1};
        bool operator==(const GraphSignature &o) const {
            return num_particles == o.num_particles && voxels == o.voxels && precond_mode == o.precond_mode && max_it[0] == o.max_it[0] &&
                   max_it[1] == o.max_it[1] && freq[0] == o.freq[0] && freq[1] == o.freq[1];
        }
    };
    GraphSignature graph_signature_;
    // key: [position buffer][binning step?][velocity-row buffer (slab migration)][halo receive buffer (slab exchanges)]
    cudaGraphExec_t graph_exec_[2][2][2][2] = {};
    uint64_t graph_kernel_nodes_[2][2][2][2] = {}; // kernels inside each graph (for blub_kernel_launch_count)
    int row_parity_ = 0;                           // toggled by every slab migration (row_ <-> row_alt_)
    struct GraphRolesAfter {
        int cur, row_parity;
        uint32_t exchanges;
    };
    GraphRolesAfter graph_roles_after_[2][2][2][2] = {};
    bool capturing_ = false;

    GridDim grid_;
    int device_;
    cudaStream_t stream_;
    bool owns_stream_;
    uint32_t max_num_particles_, num_particles_ = 0;
    float gravity_[3] = {0, 0, 0};
    uint32_t step_counter_ = 0;
    DynamicSettings dynamic_settings_;

    float4 *pos_[2] = {nullptr, nullptr}; // ping-pong for the binning reorder
    int cur_ = 0;
    float4 *row_[3] = {nullptr, nullptr, nullptr};
    GridArray<float> u_[3], density_;
    GridArray<float2> numw_[3];        // scatter-form P2G accumulators: (sum w*value, sum w) per face (sharded step / comparison path only)
    FluidBits fluid_bits_ = {nullptr, 0}; // 1 bit per cell "FLUID", rebuilt with every finished marker volume
    bool fluid_bits_stale_ = false;       // the marker volume was written through a tap since
    GridArray<int8_t> marker_;
    CellLists lists_ = {nullptr, nullptr, nullptr, nullptr, nullptr, {nullptr, nullptr, nullptr, nullptr}}; // per-step cell lists (P2G gather, binning)
    uint32_t *particle_words_ = nullptr;  // 1 bit per cell "a particle marked this cell" (scatter form: which accumulators to visit and re-zero)
    bool use_scatter_ = true;             // scatter form of P2G (default; z-slab ranks always); false = deterministic gather over cell lists
    const uint2 *voxels_ = nullptr;

    std::unique_ptr<PressureSolver> solver_;
    std::unique_ptr<PressureField> field_velocity_, field_density_;

    int slab_rank_ = 0, slab_world_ = 1;
    void *window_ = nullptr;
    size_t window_bytes_ = 0;

    StepParams *params_dev_ = nullptr;
    StepParams *params_host_ = nullptr; // pinned ring of 64
    uint32_t step_param_slot_ = 0;
    cudaEvent_t param_events_[64] = {};
};

void voxelize_rigid_solid(void *rgba16f, const uint32_t dim[3], const BlubRigidObject &o, float grid_to_world_scale, const float fluid_world_position[3],
                          double total_time, double delta, int clear_first, cudaStream_t stream, BlubRigidState *state_out);

// shared animation evaluation (solids.cu) and the triangle-mesh hull voxelizer (mesh_voxelizer.cu)
void rigid_state_at(const BlubRigidObject &o, float grid_to_world_scale, const float fluid_world_position[3], double total_time, double delta,
                    BlubRigidState &out);
BlubMesh *mesh_create(const float *positions, uint32_t num_vertices, const uint32_t *indices, uint32_t num_indices, int device);
void mesh_destroy(BlubMesh *mesh);
void mesh_info(const BlubMesh &mesh, uint32_t &num_vertices, uint32_t &num_triangles);
void voxelize_mesh(void *rgba16f, const uint32_t dim[3], BlubMesh &mesh, const BlubRigidObject &o, float grid_to_world_scale, const float fluid_world_position[3],
                   double total_time, double delta, int clear_first, cudaStream_t stream, BlubRigidState *state_out);
void read_obj(const std::string &path, std::vector<float> &positions, std::vector<uint32_t> &indices);

// scene JSON (src/scene/mod.rs:19-43)
struct SceneStaticObject {
    std::string model;          // path below the reference's `models/` directory (models.rs:253)
    BlubRigidObject placement;  // pose + animation; shape = 2 (mesh)
};
struct SceneBox {
    float min[3], max[3];
};
struct SceneConfig {
    float gravity[3] = {0, 0, 0};
    float world_position[3] = {0, 0, 0};
    float grid_to_world_scale = 1.0f;
    uint32_t grid_dimension[3] = {0, 0, 0};
    uint32_t max_num_particles = 0;
    std::vector<SceneBox> fluid_cubes;
    uint32_t num_static_objects = 0;
    std::vector<SceneStaticObject> static_objects;
};
SceneConfig parse_scene_file(const std::string &path);
// Scene::create_fluid_from_config, src/scene/mod.rs:109-144
std::unique_ptr<HybridFluid> create_fluid_from_config(const SceneConfig &cfg, int device, cudaStream_t stream);

} // namespace blub
