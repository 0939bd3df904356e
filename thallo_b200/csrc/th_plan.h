// Host side of the GN/LM + PCG solver: plan data, parameter binding, the outer loop.
// Restates the roles of reference API/src/gauss_newton.t:200-323 (SolverParameters,
// HostData, PlanData), :1128-1212 (cost, init, finalize), :1545-1785 (step),
// :1806-1862 (solver parameters), util.t:456-595 (Timer) -- as a C++ class driving
// JIT-compiled kernels through the driver API.
#pragma once
#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/Thallo.h"
#include "th_jit.h"

namespace thallo {

struct UnknownDesc { std::string name; int channels = 0; long long offset = 0; int pidx = 0; long long elements = 0; std::vector<int> dims; };
struct GroupDesc { std::string name; long long count = 0; int nterms = 0; int materialize = 0; int nnz = 0; std::vector<int> domain; std::vector<int> row_nnz; };
struct ScalarDesc { int pidx; std::string ctype; };
struct VTileDesc { int roww = 0, zoff = 0, poff = 0, bytes = 0, padl = 0, coff = -1, croww = 0, cbytes = 0; };
struct StageDesc { int slot = 0; std::string ctype; int es = 4, channels = 1, roww = 0, off = 0, bytes = 0, padl = 0, center = 0; };

struct SpaceDesc { long long elements = 0; int lanes = 1, nslots = 0; };
struct SparseEndpointDesc { int sid = 0, group = 0, slot = 0; long long count = 0, targets = 0; };
struct GroupMatDesc { int nnzp = 0, nterms = 0; };
struct SpaceCoefDesc { int space = 0, slot = 0, channels = 0; };

struct PlanDesc {
    std::string name, kind, schedule;
    bool lm = false, is_double = false, usepre = false, at_output = false, gather = false;
    // gather schedule ("space" / "sep" / "gmat" lines): index spaces of the unknowns, sparse endpoints, stored-J layout
    std::vector<SpaceDesc> spaces;
    std::vector<SparseEndpointDesc> seps;
    std::vector<GroupMatDesc> gmats;
    std::vector<SpaceCoefDesc> scoefs;      // plan-owned per-space coefficient images ("scoef" lines; ptr_pidx = -(2 + space))
    std::vector<long long> dims;
    long long nunk = 0;
    std::vector<int> ptr_pidx;
    std::vector<ScalarDesc> scalars;
    std::vector<UnknownDesc> unknowns;
    std::vector<GroupDesc> groups;
    int U = 0;
    std::vector<long long> uw_dims;
    // tiled operator kernel (front end "tile"/"vtile"/"stage" lines) and hoisted-invariant image
    int ncoef = 0;
    bool tiled = false;
    int tile[3] = {1, 1, 1}, halo[3] = {0, 0, 0};
    std::vector<std::pair<long long, int>> computed;   // ComputedArrays: (elements, gradient channels); ptr_pidx = -(100 + 2k [+ 1 for the gradient image])
    int smem_bytes = 0;
    int jp_bytes = 0;                       // two-phase tile operator: shared-memory J p planes behind the pipeline stages
    int pipe = 2;                           // shared-memory pipeline stages of th_pcg_a (TMA variant)
    std::vector<VTileDesc> vtiles;
    std::vector<StageDesc> stages;
    // multi-GPU slab partition ("partition" line): ghost layers of the slowest axis included in dims
    bool multi = false;
    int ghost_lo = 0, ghost_hi = 0;
    // multi-GPU graph partition ("gpartition" line, gather schedule): dimension of the vertex domain every
    // unknown image lives on, its local extent and the ghost vertices in front of / behind the owned range
    bool gmulti = false;
    std::vector<std::pair<long long, long long>> replicated;    // flat (offset, count) of the replicated unknown images
    int part_dim = -1;
    long long part_extent = 0;
};
bool parse_descriptor(const std::string& text, PlanDesc& d, std::string& err);
int nccl_unique_id(void* out, int capacity);

struct SolverParameters {   // gauss_newton.t:200-216, defaults :41-55
    float min_relative_decrease = 1e-3f;
    float min_trust_region_radius = 1e-32f;
    float max_trust_region_radius = 1e16f;
    float q_tolerance = 0.0001f;
    float function_tolerance = 0.000001f;
    float trust_region_radius = 1e4f;
    float radius_decrease_factor = 2.0f;
    float min_lm_diagonal = 1e-6f;
    float max_lm_diagonal = 1e32f;
    float max_solver_time_in_seconds = 0.f;
    int residual_reset_period = 10;
    int nIter = 0;
    int nIterations = 10;
    int lIterations = 10;
};

struct StateOptions {
    Thallo_InitializationParameters init{};
    cudaStream_t stream = nullptr;
};

// mirrors the device structs in skeleton/thallo_prelude.cuh
struct HScalars {
    double rz[2], aD, q, Q0, cost, modelcost, spare, red[2];
    unsigned int ticket[8];
    int it, done, lin_done, pad;
};
struct HHostFlags { volatile long long progress; volatile long long exit_word; };
// multi-GPU peer memory (mirrors ThMail / ThPeers in skeleton/thallo_prelude.cuh)
constexpr int kMaxRanks = 16, kMailKinds = 4;
struct HMail { unsigned long long q[4]; };
constexpr size_t kMailBytes = (size_t)kMailKinds * 2 * kMaxRanks * sizeof(HMail);
struct HPeers { void* box[kMaxRanks]; int rank, world, fused, epoch; };

class Plan {
public:
    Plan(const StateOptions* opts, const PlanDesc& desc, const std::string& source);
    ~Plan();
    bool ok() const { return ok_; }
    const std::string& error() const { return error_; }

    void init(void** params);
    int step(void** params);
    void solve(void** params);
    double cost();
    void set_parameter(const char* name, const void* value);
    void get_parameter(const char* name, void* value);
    void summary(Thallo_PerformanceSummary* s) const { *s = perf_; }
    long long read_vector(const char* name, void* dst, long long count);
    void* vector_pointer(const char* name);
    long long export_jacobian(int group, void* host_vals, long long* host_cols, long long capacity);
    // multi-GPU (include/thallo_b200.h "slab partition")
    int comm_init(const void* nccl_id, int rank, int world);
    int ipc_handle(void* handle64, long long* slow_extent);
    int connect(const void* handle_lo, long long extent_lo, const void* handle_hi, long long extent_hi);
    void allreduce_vec(int vec, long long offset, long long count);
    int connect_graph(const void* handle_lo, long long extent_lo, long long width_lo, const void* handle_hi, long long extent_hi,
                      long long width_hi);
    // all-to-all peer mapping (mailboxes of the in-kernel all-reduce + the neighbours' ghost layers): info = per rank
    // {local extent of the partitioned axis, bytes between consecutive solver vectors, ghost_lo, ghost_hi}
    int peer_info(long long* info4);
    int connect_all(int world, const void* handles64, const long long* infos4);
    // per-kernel device times (timingLevel >= 2, like util.t:774-790): "name count total_ms\n" lines
    std::string kernel_times();

    unsigned long long launches = 0;
    int last_linear_iterations = 0;
    unsigned long long total_linear_iterations = 0;
    const PlanDesc& desc() const { return d_; }

private:
    struct Fn { CUfunction f = nullptr; };
    CUfunction fn(const std::string& name);
    void launch(CUfunction f, dim3 grid, dim3 block, void** args, unsigned smem = 0);
    void launch_tiled(int mode);
    bool encode_map(void* dst, const void* base, int es, const std::string& ctype, int channels, int roww, bool center);
    void build_vector_maps();
    void launch_flat(CUfunction f, void** args);
    void launch_uw(CUfunction f, void** args);
    void launch_group(CUfunction f, int g, void** args);
    void bind(void** params);
    void write_lm_params();
    double compute_cost();
    void read_scalars();
    void finalize();
    void clear(void* p);
    void linear_iteration(int l);
    void rec(cudaEvent_t& e);
    void log(const char* fmt, ...) const;

    const StateOptions* opts_;
    PlanDesc d_;
    bool ok_ = false;
    std::string error_;
    CUmodule module_ = nullptr;
    std::map<std::string, CUfunction> fns_;
    size_t real_size_ = 4;
    // All solver work runs on a private stream (the legacy default stream the reference uses, util.t:769-772, cannot be
    // captured into a CUDA graph), ordered with the caller's stream by events at every API entry and exit: the caller
    // sees the same stream-ordered behaviour as with the reference (SURVEY 8b "Threading / streams").
    cudaStream_t caller() const { return opts_->stream; }
    cudaStream_t stream() const { return work_ ? work_ : opts_->stream; }
    cudaStream_t work_ = nullptr;
    cudaEvent_t ev_enter_ = nullptr, ev_leave_ = nullptr;
    int entered_ = 0;
    void enter();
    void leave();
    struct Scope { Plan* p; explicit Scope(Plan* q) : p(q) { p->enter(); } ~Scope() { p->leave(); } };
    // CUDA graphs of chunks of PCG iterations (the kernels take no per-step argument), keyed by the reset pattern of
    // the chunk and the kernel-argument images
    struct GraphEntry { cudaGraphExec_t exec = nullptr; unsigned long long launches = 0; };
    std::map<std::string, GraphEntry> graphs_;
    bool graphs_enabled_ = true;
    int graph_chunk_ = 10;
    bool use_graphs() const;
    void run_chunk(int l0, int n);

    // device state
    char* vec_block_ = nullptr;
    size_t vec_stride_ = 0;
    void* vecs_[13] = {};          // delta r b Adelta z p Ap CtC pre SSq prevX initX p2
    void* coef_ = nullptr;         // hoisted PCG-invariant coefficients, ncoef reals per element
    std::vector<char> maps_buf_;   // host image of the device struct ThMaps (128-byte CUtensorMap each)
    bool use_tma_ = false;
    bool vector_maps_ok_ = false;
    CUfunction pcg_a_ = nullptr;
    // multi-GPU state
    void* comm_ = nullptr;              // ncclComm_t
    int rank_ = 0, world_ = 1;
    char* peer_[2] = {nullptr, nullptr};            // neighbours' solver-vector blocks (CUDA IPC mappings): lo, hi
    long long peer_extent_[2] = {0, 0};
    long long peer_width_[2] = {0, 0};              // graph partition: width of the neighbour's ghost block this rank fills
    void allreduce(size_t scalars_offset, int count);
    void halo_push(int vec, int check_done);
    struct Seg { const void* src; void* dst; long long lo, count; };
    int segments(int vec, Seg* out) const;          // boundary layers of solver vector `vec` -> the neighbours' ghost layers
    void build_push(int vec, std::vector<char>& image) const;
    char* peer_all_[kMaxRanks] = {};                // every rank's solver-vector block (CUDA IPC mappings; own = vec_block_)
    size_t peer_stride_[kMaxRanks] = {};
    bool push_kernel_ = true;
    unsigned push_grid_ = 1;                        // CTAs of th_push_close
    bool fused_ = false;                            // PCG scalars all-reduced inside the kernels over peer memory (default once connect_all ran)
    HPeers peers_{};
    std::vector<char> push_init_, push_iter_, push_none_;   // ThPush images: PCGInit (z tiled / p gather), every iteration (z), none
    void* dscalar(size_t off) const { return (char*)d_scalars_ + off; }
    // gather schedule state: adjacency lists of the sparse endpoints (rebuilt when the caller's index array
    // changes), stored partial derivatives and J p of the materialised groups
    struct Adjacency { const void* src = nullptr; unsigned long long checksum = 0; bool valid = false; int* ptr = nullptr; int* perm = nullptr; };
    std::vector<Adjacency> adj_;
    std::vector<void*> jvals_, jp_, scoef_;
    int gather_persistent_ = 8;       // th_gather_s<i>: persistent grid of SMs x this many blocks (grid-stride loop, one reduction per block)
    bool gather_jtf_ = true;          // PCGInit1 gathered over the adjacency lists (THALLO_B200_SCATTER_JTF=1: the atomic scatter form)
    std::vector<void*> computed_;  // value image, gradient image per ComputedArray (2 entries each)
    void run_precompute();         // gpu.precompute, gauss_newton.t:979-986
    std::vector<char> gather_buf_;          // host image of the device struct ThGather
    unsigned long long* d_checksum_ = nullptr;
    void build_adjacency(bool verify_contents);
    void launch_gather(int which);
    unsigned tiled_grid_[2] = {1, 1};   // persistent grid of th_pcg_a_ld / th_pcg_a
    unsigned tiled_smem_[2] = {0, 0};
    int sms_ = 148;
    void* d_scalars_ = nullptr;
    double* d_partials_ = nullptr;
    HScalars* h_scalars_ = nullptr;    // pinned
    HHostFlags* h_flags_ = nullptr;    // pinned + mapped
    void* d_flags_ = nullptr;
    std::vector<char> params_buf_;
    std::vector<char> vecs_buf_;
    unsigned int flat_grid_ = 1;
    int epoch_ = 0;

    // host state (HostData)
    SolverParameters sp_;
    double radius_ = 1e4, decrease_factor_ = 2.0;   // kept in the plan's precision via round()
    double prev_cost_ = 0;
    bool finalized_ = false, initialized_ = false;
    Thallo_PerformanceSummary perf_{};
    std::chrono::steady_clock::time_point t_start_;

    // tag >= 0: launched for PCG iteration `tag` of nonlinear step `epoch`.  The host issues up to `depth` iterations
    // ahead of the device's LM exit decision; launches for iterations the device had already left return at once and
    // are kept out of the per-kernel statistics (they would deflate the average launch time).
    struct Span { cudaEvent_t a = nullptr, b = nullptr; int tag = -1, epoch = 0; };
    int cur_tag_ = -1;
    std::map<int, int> epoch_lin_done_;
    struct KernelStat { std::string name; unsigned long long count = 0; double ms = 0; std::vector<Span> pending; };
    std::map<CUfunction, int> kstat_index_;
    std::vector<KernelStat> kstats_;
    std::vector<Span> event_pool_;
    void resolve_kernel_events();
    std::vector<Span> ev_total_, ev_iter_, ev_setup_, ev_linear_, ev_finish_;
    Span cur_total_, cur_iter_, cur_phase_;
    void span_begin(Span& s);
    void span_end(Span& s, std::vector<Span>& into);
    void evaluate_timers();
    double round_real(double v) const { return d_.is_double ? v : (double)(float)v; }
};

}  // namespace thallo
