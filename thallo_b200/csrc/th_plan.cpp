#include "th_plan.h"

#include <dlfcn.h>
#include <sched.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace thallo {

// ------------------------------------------------------------------ errors (exit like the reference, cuda_util.t:103-118)
static void fatal_cuda(cudaError_t e, const char* what) {
    fprintf(stderr, "thallo_b200: CUDA error %d (%s) in %s\n", (int)e, cudaGetErrorString(e), what);
    exit((int)e ? (int)e : 1);
}
#define CD(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) fatal_cuda(_e, #x); } while (0)
static void fatal_cu(CUresult r, const char* what) {
    const char* s = nullptr;
    if (DriverApi::get().GetErrorString) DriverApi::get().GetErrorString(r, &s);
    fprintf(stderr, "thallo_b200: driver error %d (%s) in %s\n", (int)r, s ? s : "?", what);
    exit(1);
}
#define CU(x) do { CUresult _r = (x); if (_r != CUDA_SUCCESS) fatal_cu(_r, #x); } while (0)

// ------------------------------------------------------------------ descriptor
bool parse_descriptor(const std::string& text, PlanDesc& d, std::string& err) {
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        std::istringstream ls(line);
        std::string key;
        ls >> key;
        if (key == "name") ls >> d.name;
        else if (key == "kind") ls >> d.kind;
        else if (key == "lm") { int v; ls >> v; d.lm = v != 0; }
        else if (key == "real") { std::string v; ls >> v; d.is_double = (v == "double"); }
        else if (key == "usepreconditioner") { int v; ls >> v; d.usepre = v != 0; }
        else if (key == "schedule") { ls >> d.schedule; d.at_output = (d.schedule == "at_output"); d.gather = (d.schedule == "gather"); }
        else if (key == "space") { SpaceDesc sp; ls >> sp.elements >> sp.lanes >> sp.nslots; d.spaces.push_back(sp); }
        else if (key == "sep") { SparseEndpointDesc e; ls >> e.sid >> e.group >> e.slot >> e.count >> e.targets; d.seps.push_back(e); }
        else if (key == "scoef") { SpaceCoefDesc c; ls >> c.space >> c.slot >> c.channels; d.scoefs.push_back(c); }
        else if (key == "gmat") { int gi; GroupMatDesc g; ls >> gi >> g.nnzp >> g.nterms; d.gmats.push_back(g); }
        else if (key == "dims") { int n; ls >> n; d.dims.resize(n); for (auto& x : d.dims) ls >> x; }
        else if (key == "nunk") ls >> d.nunk;
        else if (key == "ptrs") { int n; ls >> n; d.ptr_pidx.resize(n); for (auto& x : d.ptr_pidx) ls >> x; }
        else if (key == "scalars") {
            int n; ls >> n;
            for (int i = 0; i < n; ++i) {
                std::string tok; ls >> tok;
                size_t c = tok.find(':');
                if (c == std::string::npos) { err = "bad scalar token " + tok; return false; }
                d.scalars.push_back({atoi(tok.substr(0, c).c_str()), tok.substr(c + 1)});
            }
        } else if (key == "unknown") {
            UnknownDesc u; int nd;
            ls >> u.name >> u.channels >> u.offset >> u.pidx >> u.elements >> nd;
            u.dims.resize(nd); for (auto& x : u.dims) ls >> x;
            d.unknowns.push_back(u);
        } else if (key == "group") {
            GroupDesc g; int nd;
            ls >> g.name >> g.count >> g.nterms >> g.materialize >> g.nnz >> nd;
            g.domain.resize(nd); for (auto& x : g.domain) ls >> x;
            std::string bar; ls >> bar;
            int v; while (ls >> v) g.row_nnz.push_back(v);
            d.groups.push_back(g);
        } else if (key == "U") ls >> d.U;
        else if (key == "uw_dims") { int n; ls >> n; d.uw_dims.resize(n); for (auto& x : d.uw_dims) ls >> x; }
        else if (key == "ncoef") ls >> d.ncoef;
        else if (key == "computed") { int k; long long n; int g; ls >> k >> n >> g; d.computed.push_back({n, g}); }
        else if (key == "partition") { ls >> d.ghost_lo >> d.ghost_hi; d.multi = true; }
        else if (key == "replicated") { long long o, n; ls >> o >> n; d.replicated.push_back({o, n}); }
        else if (key == "gpartition") { ls >> d.part_dim >> d.part_extent >> d.ghost_lo >> d.ghost_hi; d.multi = d.gmulti = true; }
        else if (key == "tile") {
            ls >> d.tile[0] >> d.tile[1] >> d.tile[2] >> d.halo[0] >> d.halo[1] >> d.halo[2] >> d.smem_bytes;
            if (!(ls >> d.pipe)) d.pipe = 2;
            d.tiled = true;
        } else if (key == "jpbytes") { ls >> d.jp_bytes;
        } else if (key == "vtile") { VTileDesc v; ls >> v.roww >> v.zoff >> v.poff >> v.bytes >> v.padl >> v.coff >> v.croww >> v.cbytes; d.vtiles.push_back(v); }
        else if (key == "stage") { StageDesc t; ls >> t.slot >> t.ctype >> t.es >> t.channels >> t.roww >> t.off >> t.bytes >> t.padl >> t.center; d.stages.push_back(t); }
        if (ls.fail() && !ls.eof()) { err = "malformed descriptor line: " + line; return false; }
    }
    if (d.nunk <= 0 || d.unknowns.empty() || d.groups.empty()) { err = "descriptor lacks unknowns or residual groups"; return false; }
    if (d.kind != "gauss_newton" && d.kind != "levenberg_marquardt") { err = "solver kind must be gauss_newton or levenberg_marquardt"; return false; }
    return true;
}

static const int kNumVecs = 13;
static const char* kVecNames[kNumVecs] = {"delta", "r", "b", "Adelta", "z", "p", "Ap_X", "CtC", "preconditioner", "SSq", "prevX", "initX", "p2"};
enum { V_DELTA, V_R, V_B, V_ADELTA, V_Z, V_P, V_AP, V_CTC, V_PRE, V_SSQ, V_PREVX, V_INITX, V_P2 };

// ------------------------------------------------------------------ NCCL (loaded at run time; only multi-GPU plans need it)
struct NcclId { char internal[128]; };
struct NcclApi {
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = nullptr;
    if (const char* e = getenv("THALLO_B200_NCCL")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    for (const char* name : {"libnccl.so.2", "libnccl.so"})      // already mapped when the host process imported torch
        if (!h) h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
    return api;
}
int nccl_unique_id(void* out, int capacity) {
    if (capacity < (int)sizeof(NcclId) || !nccl().ok) return 1;
    return nccl().GetUniqueId((NcclId*)out);
}
static void fatal_nccl(int r, const char* what) {
    fprintf(stderr, "thallo_b200: NCCL error %d (%s) in %s\n", r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", what);
    exit(1);
}

int Plan::comm_init(const void* id, int rank, int world) {
    if (!d_.multi) { error_ = "plan was not lowered with a partition"; return 1; }
    if (!nccl().ok) { error_ = "libnccl.so.2 not found (set THALLO_B200_NCCL)"; return 1; }
    NcclId nid;
    memcpy(&nid, id, sizeof nid);
    rank_ = rank; world_ = world;
    const int r = nccl().CommInitRank(&comm_, world, nid, rank);
    if (r != 0) fatal_nccl(r, "ncclCommInitRank");
    return 0;
}
int Plan::ipc_handle(void* handle64, long long* slow_extent) {
    cudaIpcMemHandle_t h;
    CD(cudaIpcGetMemHandle(&h, vec_block_));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    if (slow_extent) *slow_extent = d_.gmulti ? d_.part_extent : d_.uw_dims.back();
    return 0;
}
int Plan::connect(const void* handle_lo, long long extent_lo, const void* handle_hi, long long extent_hi) {
    const void* hs[2] = {handle_lo, handle_hi};
    const long long ex[2] = {extent_lo, extent_hi};
    for (int i = 0; i < 2; ++i) {
        if (!hs[i]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs[i], 64);
        CD(cudaIpcOpenMemHandle((void**)&peer_[i], h, cudaIpcMemLazyEnablePeerAccess));
        peer_extent_[i] = ex[i];
    }
    return 0;
}
// sum a segment of a solver vector over all ranks, in place (replicated unknowns: the camera block of bundle adjustment)
void Plan::allreduce_vec(int vec, long long offset, long long count) {
    if (!comm_) { fprintf(stderr, "thallo_b200: multi-GPU plan used before ThalloB200_PlanInitComm\n"); exit(1); }
    void* ptr = (char*)vecs_[vec] + (size_t)offset * real_size_;
    const int r = nccl().AllReduce(ptr, ptr, (size_t)count, d_.is_double ? /*ncclFloat64*/ 8 : /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm_, stream());
    if (r != 0) fatal_nccl(r, "ncclAllReduce (vector segment)");
}
// graph partition: besides the mapping, the width of each neighbour's ghost block that faces this rank
int Plan::connect_graph(const void* handle_lo, long long extent_lo, long long width_lo, const void* handle_hi, long long extent_hi,
                        long long width_hi) {
    if (!d_.gmulti) { error_ = "plan was not lowered with a graph partition"; return 1; }
    peer_width_[0] = width_lo; peer_width_[1] = width_hi;
    const long long owned = d_.part_extent - d_.ghost_lo - d_.ghost_hi;
    if (width_lo > owned || width_hi > owned) { error_ = "a neighbour's ghost block is wider than this rank's owned range"; return 1; }
    return connect(handle_lo, extent_lo, handle_hi, extent_hi);
}
// sum a few doubles of the device scalar block over all ranks, in place, on the solver stream (NVLink / NVLS)
void Plan::allreduce(size_t off, int count) {
    if (!comm_) { fprintf(stderr, "thallo_b200: multi-GPU plan used before ThalloB200_PlanInitComm\n"); exit(1); }
    const int r = nccl().AllReduce(dscalar(off), dscalar(off), (size_t)count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm_, stream());
    if (r != 0) fatal_nccl(r, "ncclAllReduce");
}
// Boundary layers of solver vector `vec` -> the neighbours' ghost layers, stored directly into their
// memory over NVLink (CUDA IPC peer mappings).  Ordering: the push precedes this rank's next
// all-reduce contribution in stream order, and a neighbour reads its ghost layers only after that
// all-reduce has completed on its side.
int Plan::segments(int vec, Seg* out) const {
    // slab: D layers of the slowest axis, h stencil-halo layers per side; graph: D local vertices, per side as many
    // vertices as the neighbour keeps ghosts of this rank's
    if (!d_.multi || world_ == 1 || (d_.uw_dims.empty() && !d_.gmulti)) return 0;
    const int nd = (int)d_.uw_dims.size();
    if (d_.gmulti && d_.part_dim < 0) return 0;       // no ghosted unknown domain (replicated / block-local unknowns only)
    const long long D = d_.gmulti ? d_.part_extent : d_.uw_dims[nd - 1];
    const long long hs[2] = {d_.gmulti ? peer_width_[0] : (long long)d_.halo[nd - 1], d_.gmulti ? peer_width_[1] : (long long)d_.halo[nd - 1]};
    int n = 0;
    for (int side = 0; side < 2; ++side) {
        if (!peer_[side] || hs[side] == 0) continue;
        const long long h = hs[side];
        const long long E = peer_extent_[side];
        long long peer_off = 0, peer_nunk = 0;
        for (auto& u : d_.unknowns) peer_nunk += (u.elements / D) * E * u.channels;
        const size_t peer_stride = ((size_t)peer_nunk * real_size_ + 255) / 256 * 256;
        for (auto& u : d_.unknowns) {
            const long long layer = (u.elements / D) * u.channels;           // scalars per slow-axis layer
            const long long src_row = side == 0 ? d_.ghost_lo : D - d_.ghost_hi - h;
            const long long dst_row = side == 0 ? E - h : 0;
            out[n].lo = u.offset + src_row * layer;
            out[n].count = (long long)h * layer;
            out[n].src = (const char*)vecs_[vec] + (size_t)out[n].lo * real_size_;
            out[n].dst = peer_[side] + peer_stride * vec + (size_t)(peer_off + dst_row * layer) * real_size_;
            peer_off += (u.elements / D) * E * u.channels;
            ++n;
        }
    }
    return n;
}
void Plan::halo_push(int vec, int check_done) {
    Seg g[8];
    if (d_.unknowns.size() > 4) { fprintf(stderr, "thallo_b200: more than 4 unknown images in a partitioned plan\n"); exit(1); }
    int n = segments(vec, g);
    if (!n) return;
    // ThSegs: src[2*NU], dst[2*NU], count[2*NU]
    const size_t nu2 = 2 * d_.unknowns.size();
    std::vector<char> buf(nu2 * 24, 0);
    for (int i = 0; i < n; ++i) {
        memcpy(buf.data() + 8 * i, &g[i].src, 8);
        memcpy(buf.data() + 8 * (nu2 + i), &g[i].dst, 8);
        memcpy(buf.data() + 8 * (2 * nu2 + i), &g[i].count, 8);
    }
    void* a[] = {buf.data(), &n, &d_scalars_, &check_done};
    launch(fn("th_halo_push"), dim3(32), dim3(256), a);
}
// ThPush image (skeleton/thallo_prelude.cuh): lo[2 NU], hi[2 NU], dst[2 NU], vec4[2 NU], n, pad
void Plan::build_push(int vec, std::vector<char>& image) const {
    const size_t nu2 = 2 * d_.unknowns.size();
    image.assign(nu2 * 28 + 8, 0);
    if (vec < 0 || !fused_) return;
    Seg g[8];
    const int n = segments(vec, g);
    for (int i = 0; i < n; ++i) {
        const long long hi = g[i].lo + g[i].count;
        const int v4 = (g[i].lo % 4 == 0 && hi % 4 == 0 && (reinterpret_cast<uintptr_t>(g[i].dst) % (4 * real_size_)) == 0) ? 1 : 0;
        memcpy(image.data() + 8 * i, &g[i].lo, 8);
        memcpy(image.data() + 8 * (nu2 + i), &hi, 8);
        memcpy(image.data() + 8 * (2 * nu2 + i), &g[i].dst, 8);
        memcpy(image.data() + 24 * nu2 + 4 * i, &v4, 4);
    }
    memcpy(image.data() + 28 * nu2, &n, 4);
}
int Plan::peer_info(long long* info) {
    info[0] = d_.gmulti ? d_.part_extent : (d_.uw_dims.empty() ? 0 : d_.uw_dims.back());
    info[1] = (long long)vec_stride_;
    info[2] = d_.ghost_lo;
    info[3] = d_.ghost_hi;
    return 0;
}
// Map every rank's solver-vector block (its tail holds the rank's mailboxes).  From then on the PCG scalars are
// all-reduced inside the kernels that produce them and boundary layers are pushed by the kernels that compute them
// (THALLO_B200_MG_NCCL=1 keeps the NCCL all-reduce + th_halo_push + th_mg_close sequence for comparison).
int Plan::connect_all(int world, const void* handles64, const long long* infos4) {
    if (!d_.multi) { error_ = "plan was not lowered with a partition"; return 1; }
    if (world != world_ || world > kMaxRanks) { error_ = "connect_all: world size does not match the communicator (or exceeds 16)"; return 1; }
    for (int r = 0; r < world; ++r) {
        peer_stride_[r] = (size_t)infos4[4 * r + 1];
        if (r == rank_) { peer_all_[r] = vec_block_; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles64 + 64 * r, 64);
        CD(cudaIpcOpenMemHandle((void**)&peer_all_[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    for (int side = 0; side < 2; ++side) {
        const int r = rank_ + (side ? 1 : -1);
        if (r < 0 || r >= world) continue;
        peer_[side] = peer_all_[r];
        peer_extent_[side] = infos4[4 * r];
        peer_width_[side] = side ? infos4[4 * r + 2] : infos4[4 * r + 3];     // the neighbour's ghost block facing this rank
    }
    if (d_.gmulti && d_.part_dim >= 0) {
        const long long owned = d_.part_extent - d_.ghost_lo - d_.ghost_hi;
        if (peer_width_[0] > owned || peer_width_[1] > owned) { error_ = "a neighbour's ghost block is wider than this rank's owned range"; return 1; }
    }
    const char* e = getenv("THALLO_B200_MG_NCCL");
    fused_ = !(e && atoi(e) != 0);
    peers_ = HPeers{};
    for (int r = 0; r < world; ++r) peers_.box[r] = peer_all_[r] + peer_stride_[r] * kNumVecs;
    long long push_bytes = 0;
    {
        Seg g[8];
        const int n = fused_ ? segments(V_Z, g) : 0;
        for (int i = 0; i < n; ++i) push_bytes += g[i].count * (long long)real_size_;
    }
    // fused: 1 = the boundary layers of z are pushed by the last CTA of th_pcg_b, 2 = by the small th_push_close kernel.
    // Measured at 8 GPUs (profiles/r02h_n8_*): a single CTA needs ~100 us for the 1.2 MB faces of a 160^3 slab (the
    // kernel: 10.5 k instead of 6.4 k iterations/s), while for the 48-128 KB rows of the 2-D slabs and the graph the extra
    // launch costs 1-6 %.  THALLO_B200_PUSH_KERNEL=0/1 forces one or the other.
    push_kernel_ = push_bytes > 256 * 1024;
    if (const char* pk = getenv("THALLO_B200_PUSH_KERNEL")) push_kernel_ = atoi(pk) != 0;
    peers_.rank = rank_; peers_.world = world_; peers_.fused = fused_ ? (push_kernel_ ? 2 : 1) : 0;
    build_push(d_.tiled ? V_Z : V_P, push_init_);
    build_push(V_Z, push_iter_);
    push_grid_ = (unsigned)std::max<long long>(1, std::min<long long>(64, (push_bytes + 8191) / 8192));      // CTAs of th_push_close
    return 0;
}

// ------------------------------------------------------------------ plan

void Plan::log(const char* fmt, ...) const {
    if (opts_->init.verbosityLevel <= 0) return;
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
}

Plan::Plan(const StateOptions* opts, const PlanDesc& desc, const std::string& source) : opts_(opts), d_(desc) {
    real_size_ = d_.is_double ? 8 : 4;
    const DriverApi& api = DriverApi::get();
    if (!api.ok) { error_ = "no CUDA device / driver available: thallo_b200 has no CPU fallback"; return; }
    std::vector<char> cubin;
    std::string clog;
    std::vector<std::string> xopts;           // tuning experiments: extra NVRTC options, e.g. "-DTH_OWN_ENDPOINT=0"
    if (const char* e = getenv("THALLO_B200_NVRTC_OPTS")) {
        std::stringstream ss(e);
        std::string o;
        while (ss >> o) xopts.push_back(o);
    }
    if (!compile_cubin(source, skeleton_dir(), cubin, clog, xopts)) { error_ = "NVRTC compilation failed:\n" + clog; return; }
    if (opts_->init.verbosityLevel > 1 && !clog.empty()) printf("%s\n", clog.c_str());
    CUresult r = api.ModuleLoadData(&module_, cubin.data());
    if (r != CUDA_SUCCESS) { error_ = "cuModuleLoadData failed (" + std::to_string((int)r) + ")"; return; }

    if (!getenv("THALLO_B200_NO_PRIVATE_STREAM")) {
        CD(cudaStreamCreateWithFlags(&work_, cudaStreamNonBlocking));
        CD(cudaEventCreateWithFlags(&ev_enter_, cudaEventDisableTiming));
        CD(cudaEventCreateWithFlags(&ev_leave_, cudaEventDisableTiming));
    }
    if (const char* e = getenv("THALLO_B200_GRAPH")) graphs_enabled_ = atoi(e) != 0;
    if (const char* e = getenv("THALLO_B200_GRAPH_CHUNK")) graph_chunk_ = std::max(1, atoi(e));
    Scope scope(this);
    // solver vectors: one allocation, 12 unknown-sized vectors (gauss_newton.t:1963-2071)
    vec_stride_ = ((size_t)d_.nunk * real_size_ + 255) / 256 * 256;
    // (+ the mailboxes of the multi-GPU in-kernel all-reduce behind the last vector, so that one IPC handle covers both)
    CD(cudaMalloc((void**)&vec_block_, vec_stride_ * kNumVecs + kMailBytes));
    CD(cudaMemsetAsync(vec_block_, 0, vec_stride_ * kNumVecs + kMailBytes, stream()));
    for (int i = 0; i < kNumVecs; ++i) vecs_[i] = vec_block_ + vec_stride_ * i;
    for (auto& c : d_.computed) {     // ComputedArrays: value image + gradient image (ImageTemporary, thallo.t:1806-1815)
        void* v = nullptr; void* g = nullptr;
        CD(cudaMalloc(&v, (size_t)c.first * real_size_));
        CD(cudaMemsetAsync(v, 0, (size_t)c.first * real_size_, stream()));
        if (c.second > 0) {
            CD(cudaMalloc(&g, (size_t)c.first * c.second * real_size_));
            CD(cudaMemsetAsync(g, 0, (size_t)c.first * c.second * real_size_, stream()));
        }
        computed_.push_back(v); computed_.push_back(g);
    }
    if (d_.ncoef > 0) {
        const size_t n = (size_t)d_.unknowns[0].elements * d_.ncoef * real_size_;
        CD(cudaMalloc(&coef_, n));
        CD(cudaMemsetAsync(coef_, 0, n, stream()));
    }
    CD(cudaMalloc(&d_scalars_, sizeof(HScalars)));
    CD(cudaMemsetAsync(d_scalars_, 0, sizeof(HScalars), stream()));
    CD(cudaHostAlloc((void**)&h_scalars_, sizeof(HScalars), cudaHostAllocDefault));
    CD(cudaHostAlloc((void**)&h_flags_, sizeof(HHostFlags), cudaHostAllocMapped));
    h_flags_->progress = 0; h_flags_->exit_word = 0;
    CD(cudaHostGetDevicePointer(&d_flags_, (void*)h_flags_, 0));

    int sms = 148;
    int dev = 0;
    CD(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    sms_ = sms;
    const long long want = (d_.nunk + 255) / 256;
    flat_grid_ = (unsigned)std::max(1LL, std::min<long long>(want, (long long)sms * 8));
    // partial sums: up to 2 values per block of the largest grid
    long long maxblocks = flat_grid_;
    if (d_.at_output) {
        long long b = 1;
        if (d_.uw_dims.size() == 1) b = (d_.uw_dims[0] + 255) / 256;
        else if (d_.uw_dims.size() == 2) b = ((d_.uw_dims[0] + 31) / 32) * ((d_.uw_dims[1] + 7) / 8);
        else b = ((d_.uw_dims[0] + 7) / 8) * ((d_.uw_dims[1] + 7) / 8) * ((d_.uw_dims[2] + 3) / 4);
        maxblocks = std::max(maxblocks, b);
        if (d_.tiled) maxblocks = std::max<long long>(maxblocks, (long long)sms * 32);   // persistent grid <= SMs x resident CTAs
    }
    for (auto& g : d_.groups) maxblocks = std::max(maxblocks, (g.count + 255) / 256);
    for (auto& sp : d_.spaces) maxblocks = std::max(maxblocks, (sp.elements * sp.lanes + 255) / 256);
    CD(cudaMalloc((void**)&d_partials_, sizeof(double) * 2 * (size_t)maxblocks));
    if (d_.gather) {
        if (const char* e = getenv("THALLO_B200_SCATTER_JTF")) gather_jtf_ = atoi(e) == 0;
        if (const char* e = getenv("THALLO_B200_GATHER_BLOCKS_PER_SM")) gather_persistent_ = atoi(e);     // 0: one block per 256 threads of work
        adj_.resize(d_.seps.size());
        jvals_.assign(d_.groups.size(), nullptr);
        jp_.assign(d_.groups.size(), nullptr);
        for (size_t g = 0; g < d_.groups.size(); ++g) {
            if (!d_.groups[g].materialize) continue;          // 1: J and J p stored ([Jt][[J]p]); 2: only J p stored (Jt[Jp])
            if (d_.groups[g].materialize == 1)
                CD(cudaMalloc(&jvals_[g], (size_t)d_.groups[g].count * d_.gmats[g].nnzp * real_size_));
            CD(cudaMalloc(&jp_[g], (size_t)d_.groups[g].count * d_.gmats[g].nterms * real_size_));
        }
        scoef_.assign(d_.spaces.size(), nullptr);
        for (auto& c : d_.scoefs)
            CD(cudaMalloc(&scoef_[c.space], (size_t)d_.spaces[c.space].elements * c.channels * real_size_));
        CD(cudaMalloc((void**)&d_checksum_, 8));
        const size_t ns = std::max<size_t>(1, d_.seps.size());
        gather_buf_.assign(8 * (2 * ns + 2 * d_.groups.size()), 0);
        for (size_t g = 0; g < d_.groups.size(); ++g) {
            memcpy(gather_buf_.data() + 8 * (2 * ns + g), &jvals_[g], 8);
            memcpy(gather_buf_.data() + 8 * (2 * ns + d_.groups.size() + g), &jp_[g], 8);
        }
    }

    // kernel-argument images of the device structs
    const size_t nptr = std::max<size_t>(1, d_.ptr_pidx.size()), nsc = std::max<size_t>(1, d_.scalars.size());
    size_t psz = nptr * 8 + (nsc + 4) * real_size_;
    psz = (psz + 7) / 8 * 8;
    params_buf_.assign(psz, 0);
    vecs_buf_.assign(11 * sizeof(void*), 0);
    memcpy(vecs_buf_.data(), vecs_, 10 * sizeof(void*));
    memcpy(vecs_buf_.data() + 10 * sizeof(void*), &vecs_[V_P2], sizeof(void*));
    if (d_.tiled) {
        maps_buf_.assign(128 * (5 * d_.unknowns.size() + std::max<size_t>(1, d_.stages.size())) + 64, 0);
        build_vector_maps();
        // persistent grids: SMs x CTAs resident per SM (d_.pipe pipeline stages of shared memory with TMA, one without)
        long long ntiles = 1;
        for (size_t i = 0; i < d_.uw_dims.size(); ++i) ntiles *= (d_.uw_dims[i] + d_.tile[i] - 1) / d_.tile[i];
        const char* names[2] = {"th_pcg_a_ld", "th_pcg_a"};
        for (int v = 0; v < 2; ++v) {
            CUfunction f = fn(names[v]);
            tiled_smem_[v] = (unsigned)d_.smem_bytes * (v ? (unsigned)d_.pipe : 1u) + (unsigned)d_.jp_bytes;
            if (tiled_smem_[v] > 48 * 1024 && api.FuncSetAttribute)
                CU(api.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)tiled_smem_[v]));
            int per_sm = 2;
            const int threads = d_.tile[0] * d_.tile[1] * d_.tile[2];
            if (api.OccupancyMaxActiveBlocks && api.OccupancyMaxActiveBlocks(&per_sm, f, threads, tiled_smem_[v]) != CUDA_SUCCESS) per_sm = 2;
            per_sm = std::max(1, std::min(per_sm, 32));
            if (const char* e = getenv("THALLO_B200_CTAS_PER_SM")) per_sm = std::max(1, std::min(atoi(e), 32));   // d_partials_ holds SMs x 32 blocks
            tiled_grid_[v] = (unsigned)std::min<long long>(ntiles, (long long)sms * per_sm);
        }
    }
    peers_ = HPeers{};
    peers_.world = 1;
    build_push(-1, push_none_);
    push_init_ = push_iter_ = push_none_;
    sp_ = SolverParameters();
    ok_ = true;
}

// ------------------------------------------------------------------ TMA descriptors of the staged tiles
// One tensor map per staged array: the image as a rank-ND tensor of scalars (innermost extent
// W*channels), box = tile + halo with rows padded to 16 bytes, zero fill outside the image.
static char* maps_base(std::vector<char>& b) {
    return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(b.data()) + 63) & ~uintptr_t(63));
}
bool Plan::encode_map(void* dst, const void* base, int es, const std::string& ctype, int channels, int roww, bool center) {
    const DriverApi& api = DriverApi::get();
    if (!api.TensorMapEncodeTiled) return false;
    const int nd = (int)d_.uw_dims.size();
    cuuint64_t gdim[3] = {1, 1, 1};
    cuuint64_t gstride[2] = {0, 0};
    cuuint32_t box[3] = {1, 1, 1}, estr[3] = {1, 1, 1};
    gdim[0] = (cuuint64_t)d_.uw_dims[0] * channels;
    for (int i = 1; i < nd; ++i) gdim[i] = (cuuint64_t)d_.uw_dims[i];
    gstride[0] = gdim[0] * es;
    gstride[1] = gstride[0] * gdim[1];
    box[0] = (cuuint32_t)roww;
    for (int i = 1; i < nd; ++i) box[i] = (cuuint32_t)(d_.tile[i] + (center ? 0 : 2 * d_.halo[i]));
    // the innermost start coordinate of every box (tile origin * channels - padl) must be 16-byte aligned
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (gstride[0] & 15) || box[0] > 256 || (((size_t)d_.tile[0] * channels * es) & 15)) return false;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (ctype == "uchar") dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
    else if (ctype == "int") dt = CU_TENSOR_MAP_DATA_TYPE_INT32;
    else if (es == 8) dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUtensorMap m;
    CUresult r = api.TensorMapEncodeTiled(&m, dt, (cuuint32_t)nd, const_cast<void*>(base), gdim, gstride, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    memcpy(dst, &m, 128);
    return true;
}
void Plan::build_vector_maps() {
    char* mb = maps_base(maps_buf_);
    const size_t nu = d_.unknowns.size();
    const int src[5] = {V_Z, V_P, V_P2, V_DELTA, V_CTC};      // ThMaps: z[nu], p[3][nu], c[nu]
    vector_maps_ok_ = true;
    for (int v = 0; v < 5; ++v)
        for (size_t k = 0; k < nu; ++k) {
            const char* base = (const char*)vecs_[src[v]] + (size_t)d_.unknowns[k].offset * real_size_;
            const bool center = v == 4;
            if (center && d_.vtiles[k].coff < 0) continue;      // CtC not staged
            if (!encode_map(mb + 128 * (v * nu + k), base, (int)real_size_, "real", d_.unknowns[k].channels,
                            center ? d_.vtiles[k].croww : d_.vtiles[k].roww, center))
                vector_maps_ok_ = false;
        }
}

void Plan::enter() {
    if (!work_ || entered_++ > 0) return;
    CD(cudaEventRecord(ev_enter_, caller()));
    CD(cudaStreamWaitEvent(work_, ev_enter_, 0));
}
void Plan::leave() {
    if (!work_ || --entered_ > 0) return;
    CD(cudaEventRecord(ev_leave_, work_));
    CD(cudaStreamWaitEvent(caller(), ev_leave_, 0));
}
bool Plan::use_graphs() const {
    // per-kernel event pairs (timingLevel >= 2) and NCCL calls inside the iteration keep the plain launch sequence
    return graphs_enabled_ && work_ && opts_->init.timingLevel < 2 && !(d_.multi && !fused_) && d_.replicated.empty();
}
// Iterations l0 .. l0+n-1 of the linear solve as one graph launch.
void Plan::run_chunk(int l0, int n) {
    std::string key;
    for (int k = 0; k < n; ++k) key.push_back(d_.lm && ((l0 + k + 1) % sp_.residual_reset_period) == 0 ? 'R' : '.');
    const size_t nptr = std::max<size_t>(1, d_.ptr_pidx.size()), nsc = std::max<size_t>(1, d_.scalars.size());
    key.append(params_buf_.data(), 8 * nptr + real_size_ * nsc);        // (the LM scalars behind them are read by PCGInit only)
    key.append(vecs_buf_.data(), vecs_buf_.size());
    key.append(gather_buf_.data(), gather_buf_.size());
    key.append((const char*)&peers_, sizeof(void*) * kMaxRanks + 3 * sizeof(int));
    key.append(push_iter_.data(), push_iter_.size());
    key.append((const char*)&sp_.q_tolerance, sizeof(float));
    key.push_back(use_tma_ ? 'T' : 'L');
    if (d_.tiled) key.append(maps_base(maps_buf_), 128 * (5 * d_.unknowns.size() + std::max<size_t>(1, d_.stages.size())));
    auto it = graphs_.find(key);
    if (it == graphs_.end()) {
        if (graphs_.size() >= 16) {                                          // callers that keep changing buffers: start over
            for (auto& g : graphs_) cudaGraphExecDestroy(g.second.exec);
            graphs_.clear();
        }
        const unsigned long long before = launches;
        cudaGraph_t graph = nullptr;
        CD(cudaStreamBeginCapture(work_, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < n; ++k) linear_iteration(l0 + k);
        CD(cudaStreamEndCapture(work_, &graph));
        GraphEntry e;
        e.launches = launches - before;
        launches = before;
        CD(cudaGraphInstantiate(&e.exec, graph, 0));
        CD(cudaGraphDestroy(graph));
        it = graphs_.emplace(key, e).first;
    }
    CD(cudaGraphLaunch(it->second.exec, work_));
    launches += it->second.launches;
}

Plan::~Plan() {
    if (work_) cudaStreamSynchronize(work_);
    for (auto& g : graphs_) cudaGraphExecDestroy(g.second.exec);
    bool all = false;
    for (int r = 0; r < kMaxRanks; ++r) if (peer_all_[r] && peer_all_[r] != vec_block_) { cudaIpcCloseMemHandle(peer_all_[r]); all = true; }
    if (!all) for (int i = 0; i < 2; ++i) if (peer_[i]) cudaIpcCloseMemHandle(peer_[i]);
    if (comm_) nccl().CommDestroy(comm_);
    for (auto& a : adj_) { if (a.ptr) cudaFree(a.ptr); if (a.perm) cudaFree(a.perm); }
    for (void* q : jvals_) if (q) cudaFree(q);
    for (void* q : jp_) if (q) cudaFree(q);
    for (void* q : scoef_) if (q) cudaFree(q);
    for (void* q : computed_) if (q) cudaFree(q);
    if (d_checksum_) cudaFree(d_checksum_);
    if (vec_block_) cudaFree(vec_block_);
    if (coef_) cudaFree(coef_);
    if (d_scalars_) cudaFree(d_scalars_);
    if (d_partials_) cudaFree(d_partials_);
    if (h_scalars_) cudaFreeHost(h_scalars_);
    if (h_flags_) cudaFreeHost((void*)h_flags_);
    for (auto* v : {&ev_total_, &ev_iter_, &ev_setup_, &ev_linear_, &ev_finish_})
        for (auto& s : *v) { if (s.a) cudaEventDestroy(s.a); if (s.b) cudaEventDestroy(s.b); }
    for (auto& k : kstats_) for (auto& s : k.pending) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto& s : event_pool_) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    if (module_ && DriverApi::get().ok) DriverApi::get().ModuleUnload(module_);
    if (ev_enter_) cudaEventDestroy(ev_enter_);
    if (ev_leave_) cudaEventDestroy(ev_leave_);
    if (work_) cudaStreamDestroy(work_);
}

CUfunction Plan::fn(const std::string& name) {
    auto it = fns_.find(name);
    if (it != fns_.end()) return it->second;
    CUfunction f = nullptr;
    CUresult r = DriverApi::get().ModuleGetFunction(&f, module_, name.c_str());
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "thallo_b200: kernel %s missing from the compiled plan\n", name.c_str());
        exit(1);
    }
    fns_[name] = f;
    kstat_index_[f] = (int)kstats_.size();
    kstats_.push_back(KernelStat());
    kstats_.back().name = name;
    return f;
}

void Plan::launch(CUfunction f, dim3 grid, dim3 block, void** args, unsigned smem) {
    const bool timed = opts_->init.timingLevel >= 2;   // per-kernel events, util.t:774-790
    Span s;
    if (timed) {
        if (!event_pool_.empty()) { s = event_pool_.back(); event_pool_.pop_back(); }
        else { CD(cudaEventCreate(&s.a)); CD(cudaEventCreate(&s.b)); }
        s.tag = cur_tag_; s.epoch = epoch_;
        CD(cudaEventRecord(s.a, stream()));
    }
    CU(DriverApi::get().LaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem, (CUstream)stream(), args, nullptr));
    ++launches;
    if (timed) {
        CD(cudaEventRecord(s.b, stream()));
        KernelStat& k = kstats_[kstat_index_[f]];
        k.pending.push_back(s);
        if (k.pending.size() >= 4096) resolve_kernel_events();
    }
}
void Plan::resolve_kernel_events() {
    CD(cudaStreamSynchronize(stream()));
    for (auto& k : kstats_) {
        std::vector<Span> keep;
        for (auto& s : k.pending) {
            if (s.tag >= 0) {
                auto it = epoch_lin_done_.find(s.epoch);
                if (it == epoch_lin_done_.end()) { keep.push_back(s); continue; }       // step still running: decide later
                if (s.tag >= it->second) { event_pool_.push_back(s); continue; }        // issued after the device's exit: a no-op
            }
            float t = 0;
            if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) { k.ms += t; ++k.count; }
            event_pool_.push_back(s);
        }
        k.pending.swap(keep);
    }
    while (epoch_lin_done_.size() > 64) epoch_lin_done_.erase(epoch_lin_done_.begin());
}
std::string Plan::kernel_times() {
    resolve_kernel_events();
    std::ostringstream o;
    o.precision(9);
    for (auto& k : kstats_) if (k.count) o << k.name << " " << k.count << " " << k.ms << "\n";
    return o.str();
}
void Plan::launch_flat(CUfunction f, void** args) { launch(f, dim3(flat_grid_), dim3(256), args); }
void Plan::launch_uw(CUfunction f, void** args) {
    const auto& u = d_.uw_dims;
    if (u.size() == 1) launch(f, dim3((unsigned)((u[0] + 255) / 256)), dim3(256), args);
    else if (u.size() == 2) launch(f, dim3((unsigned)((u[0] + 31) / 32), (unsigned)((u[1] + 7) / 8)), dim3(32, 8), args);
    else launch(f, dim3((unsigned)((u[0] + 7) / 8), (unsigned)((u[1] + 7) / 8), (unsigned)((u[2] + 3) / 4)), dim3(8, 8, 4), args);
}
void Plan::launch_group(CUfunction f, int g, void** args) {
    launch(f, dim3((unsigned)((d_.groups[g].count + 255) / 256)), dim3(256), args);
}
void Plan::clear(void* p) { CD(cudaMemsetAsync(p, 0, (size_t)d_.nunk * real_size_, stream())); }

// ------------------------------------------------------------------ gather schedule: adjacency lists
// For every sparse endpoint (residual group g reaching an unknown index space through index array
// I): the residual elements e grouped by target I[e], as CSR offsets + a stable permutation.  Built
// on the host from one copy of the index array (counting sort), when the caller's pointer changes
// or (checked at every Thallo_ProblemInit with a device-side checksum) its contents do.  An index
// array that is already sorted -- e.g. the reference's own per-vertex edge lists,
// examples/shared/ThalloGraph.h:67-79 -- needs no permutation.
void Plan::build_adjacency(bool verify_contents) {
    const size_t ns = std::max<size_t>(1, d_.seps.size());
    for (size_t i = 0; i < d_.seps.size(); ++i) {
        const SparseEndpointDesc& e = d_.seps[i];
        Adjacency& a = adj_[e.sid];
        const void* src = nullptr;
        memcpy(&src, params_buf_.data() + 8 * e.slot, 8);
        // contents are what matter: a caller that passes a copy of the same index array keeps the lists
        bool rebuild = !a.valid;
        unsigned long long sum = a.checksum;
        if (!a.valid || a.src != src || verify_contents) {
            CD(cudaMemsetAsync(d_checksum_, 0, 8, stream()));
            long long n = e.count;
            void* args[] = {&src, &n, &d_checksum_};
            launch(fn("th_index_checksum"), dim3((unsigned)std::min<long long>((n + 255) / 256, (long long)sms_ * 8)), dim3(256), args);
            CD(cudaMemcpyAsync(&sum, d_checksum_, 8, cudaMemcpyDeviceToHost, stream()));
            CD(cudaStreamSynchronize(stream()));
            rebuild = rebuild || sum != a.checksum;
            a.src = src;
        }
        if (!rebuild) continue;
        std::vector<int> idx((size_t)e.count);
        CD(cudaMemcpyAsync(idx.data(), src, sizeof(int) * (size_t)e.count, cudaMemcpyDeviceToHost, stream()));
        CD(cudaStreamSynchronize(stream()));
        std::vector<int> ptr((size_t)e.targets + 1, 0), perm((size_t)e.count);
        for (long long k = 0; k < e.count; ++k) {
            if (idx[k] < 0 || idx[k] >= e.targets) {
                fprintf(stderr, "thallo_b200: sparse index %d out of range [0, %lld) at element %lld\n", idx[k], e.targets, k);
                exit(1);
            }
            ++ptr[(size_t)idx[k] + 1];
        }
        for (long long t = 0; t < e.targets; ++t) ptr[t + 1] += ptr[t];
        std::vector<int> cur(ptr.begin(), ptr.end() - 1);
        bool identity = true;
        for (long long k = 0; k < e.count; ++k) {
            const int pos = cur[idx[k]]++;
            perm[pos] = (int)k;
            identity = identity && pos == k;
        }
        if (!a.ptr) CD(cudaMalloc((void**)&a.ptr, sizeof(int) * ((size_t)e.targets + 1)));
        CD(cudaMemcpyAsync(a.ptr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, stream()));
        if (identity) { if (a.perm) { cudaFree(a.perm); a.perm = nullptr; } }
        else {
            if (!a.perm) CD(cudaMalloc((void**)&a.perm, sizeof(int) * (size_t)e.count));
            CD(cudaMemcpyAsync(a.perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, stream()));
        }
        CD(cudaStreamSynchronize(stream()));
        a.src = src; a.checksum = sum; a.valid = true;
        log("adjacency of sparse endpoint %d: %lld residuals -> %lld unknown elements%s\n", e.sid, e.count, e.targets, identity ? " (sorted, no permutation)" : "");
    }
    for (size_t i = 0; i < d_.seps.size(); ++i) {
        const Adjacency& a = adj_[i];
        memcpy(gather_buf_.data() + 8 * i, &a.ptr, 8);
        memcpy(gather_buf_.data() + 8 * (ns + i), &a.perm, 8);
    }
}
// Ap (which = 0) or Adelta (which = 1) in the gather schedule: J p of the materialised groups, then one kernel per index space
void Plan::launch_gather(int which) {
    void* P = params_buf_.data();
    void* V = vecs_buf_.data();
    void* G = gather_buf_.data();
    if (!which)
        for (size_t g = 0; g < d_.groups.size(); ++g) {
            if (!d_.groups[g].materialize) continue;
            void* a[] = {P, V, G, &d_scalars_};
            launch_group(fn((d_.groups[g].materialize == 1 ? "th_matj_g" : "th_applyj_g") + std::to_string(g)), (int)g, a);
        }
    for (size_t s = 0; s < d_.spaces.size(); ++s) {
        int first = s == 0;
        int reduce_now = s + 1 == d_.spaces.size() && d_.replicated.empty();      // the last contribution to <p, Ap>: all-reduce in-kernel
        void* a[] = {P, V, G, &d_scalars_, &d_partials_, &which, &first, &reduce_now, &peers_};
        const long long threads = d_.spaces[s].elements * d_.spaces[s].lanes;
        long long blocks = (threads + 255) / 256;
        if (gather_persistent_) blocks = std::min<long long>(blocks, (long long)sms_ * gather_persistent_);
        launch(fn("th_gather_s" + std::to_string(s)), dim3((unsigned)blocks), dim3(256), a);
    }
    if (!d_.replicated.empty()) {      // replicated unknowns: sum the ranks' partial J^T J p, then + CtC p and the dot product
        long long n = 0;
        for (auto& r : d_.replicated) { allreduce_vec(which ? V_ADELTA : V_AP, r.first, r.second); n += r.second; }
        void* a[] = {P, V, &d_scalars_, &d_partials_, &which, &peers_};
        launch(fn("th_rep_finish"), dim3((unsigned)std::min<long long>((n + 255) / 256, (long long)sms_ * 8)), dim3(256), a);
    }
}

// util.initParameters, util.t:609-643: images / sparse = device pointers, scalars read through host pointers
void Plan::bind(void** params) {
    char* buf = params_buf_.data();
    const size_t nptr = std::max<size_t>(1, d_.ptr_pidx.size()), nsc = std::max<size_t>(1, d_.scalars.size());
    for (size_t i = 0; i < d_.ptr_pidx.size(); ++i) {
        if (d_.ptr_pidx[i] <= -100) memcpy(buf + 8 * i, &computed_[-(d_.ptr_pidx[i] + 100)], 8);   // plan-owned ComputedArray value / gradient image
        else if (d_.ptr_pidx[i] < -1) memcpy(buf + 8 * i, &scoef_[-(d_.ptr_pidx[i] + 2)], 8);   // plan-owned per-space coefficient image
        else if (d_.ptr_pidx[i] < 0) memcpy(buf + 8 * i, &coef_, 8);          // plan-owned coefficient image
        else memcpy(buf + 8 * i, &params[d_.ptr_pidx[i]], 8);
    }
    if (d_.tiled) {
        // tensor maps of the caller-owned staged images are re-encoded on every bind (the caller may
        // pass different buffers at every step, util.t:609-643); any image TMA cannot describe
        // (unaligned base or row pitch) switches the plan to the cooperative-load variant.
        char* mb = maps_base(maps_buf_);
        bool ok = vector_maps_ok_ && !getenv("THALLO_B200_NO_TMA");
        for (size_t s = 0; s < d_.stages.size() && ok; ++s) {
            const StageDesc& t = d_.stages[s];
            void* base = nullptr;
            memcpy(&base, buf + 8 * t.slot, 8);
            ok = encode_map(mb + 128 * (5 * d_.unknowns.size() + s), base, t.es, t.ctype, t.channels, t.roww, t.center != 0);
        }
        use_tma_ = ok;
        pcg_a_ = fn(ok ? "th_pcg_a" : "th_pcg_a_ld");
    }
    char* sc = buf + 8 * nptr;
    for (size_t i = 0; i < d_.scalars.size(); ++i) {
        const void* hp = params[d_.scalars[i].pidx];
        double v = 0;
        const std::string& t = d_.scalars[i].ctype;
        if (t == "float") v = *(const float*)hp;
        else if (t == "double") v = *(const double*)hp;
        else if (t == "int") v = *(const int*)hp;
        else if (t == "uchar") v = *(const unsigned char*)hp;
        if (d_.is_double) memcpy(sc + 8 * i, &v, 8);
        else { float f = (float)v; memcpy(sc + 4 * i, &f, 4); }
    }
    (void)nsc;
    write_lm_params();
    if (d_.gather) build_adjacency(false);
}
void Plan::write_lm_params() {
    const size_t nptr = std::max<size_t>(1, d_.ptr_pidx.size()), nsc = std::max<size_t>(1, d_.scalars.size());
    char* lm = params_buf_.data() + 8 * nptr + real_size_ * nsc;
    const double v[4] = {radius_, decrease_factor_, sp_.min_lm_diagonal, sp_.max_lm_diagonal};
    for (int i = 0; i < 4; ++i) {
        if (d_.is_double) memcpy(lm + 8 * i, &v[i], 8);
        else { float f = (float)v[i]; memcpy(lm + 4 * i, &f, 4); }
    }
}

void Plan::read_scalars() {
    CD(cudaMemcpyAsync(h_scalars_, d_scalars_, sizeof(HScalars), cudaMemcpyDeviceToHost, stream()));
    CD(cudaStreamSynchronize(stream()));
}

// gpu.precompute (gauss_newton.t:979-986): refresh every ComputedArray and its gradient image from the current unknowns
void Plan::run_precompute() {
    for (size_t k = 0; k < d_.computed.size(); ++k) {
        void* a[] = {params_buf_.data()};
        launch(fn("th_precompute_c" + std::to_string(k)), dim3((unsigned)((d_.computed[k].first + 255) / 256)), dim3(256), a);
    }
}

// computeCost, gauss_newton.t:1128-1136
double Plan::compute_cost() {
    void* P = params_buf_.data();
    if (d_.at_output) {            // every group lives on the unknown domain: one pass for all of them
        void* args[] = {P, &d_scalars_, &d_partials_};
        launch_uw(fn("th_cost_uw"), args);
    } else
        for (size_t g = 0; g < d_.groups.size(); ++g) {
            int first = g == 0;
            void* args[] = {P, &d_scalars_, &d_partials_, &first};
            launch_group(fn("th_cost_g" + std::to_string(g)), (int)g, args);
        }
    if (d_.multi) allreduce(offsetof(HScalars, cost), 1);
    read_scalars();
    return round_real(h_scalars_->cost);
}

void Plan::span_begin(Span& s) {
    if (!s.a) { CD(cudaEventCreate(&s.a)); CD(cudaEventCreate(&s.b)); }
    CD(cudaEventRecord(s.a, stream()));
}
void Plan::span_end(Span& s, std::vector<Span>& into) {
    CD(cudaEventRecord(s.b, stream()));
    into.push_back(s);
    s = Span();
}

static void accumulate(Thallo_PerformanceEntry& e, const std::vector<double>& ms) {
    e = Thallo_PerformanceEntry{};
    if (ms.empty()) return;
    e.count = (unsigned)ms.size();
    double mn = ms[0], mx = ms[0], sum = 0;
    for (double v : ms) { mn = std::min(mn, v); mx = std::max(mx, v); sum += v; }
    const double mean = sum / ms.size();
    double var = 0;
    for (double v : ms) var += (v - mean) * (v - mean);
    e.minMS = mn; e.maxMS = mx; e.meanMS = mean;
    e.stddevMS = ms.size() > 1 ? std::sqrt(var / (ms.size() - 1)) : 0.0;
}
// Timer:evaluate, util.t:516-541: five buckets
void Plan::evaluate_timers() {
    struct B { std::vector<Span>* v; Thallo_PerformanceEntry* e; } bs[] = {
        {&ev_total_, &perf_.total}, {&ev_iter_, &perf_.nonlinearIteration}, {&ev_setup_, &perf_.nonlinearSetup},
        {&ev_linear_, &perf_.linearSolve}, {&ev_finish_, &perf_.nonlinearResolve}};
    for (auto& b : bs) {
        std::vector<double> ms;
        for (auto& s : *b.v) {
            float t = 0;
            if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms.push_back(t);
            cudaEventDestroy(s.a); cudaEventDestroy(s.b);
        }
        b.v->clear();
        accumulate(*b.e, ms);
    }
}

// init, gauss_newton.t:1166-1198
void Plan::init(void** params) {
    Scope scope(this);
    finalized_ = false;
    initialized_ = true;
    t_start_ = std::chrono::steady_clock::now();
    for (auto* v : {&ev_total_, &ev_iter_, &ev_setup_, &ev_linear_, &ev_finish_}) {
        for (auto& s : *v) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
        v->clear();
    }
    span_begin(cur_total_);
    sp_.nIter = 0;
    radius_ = round_real(sp_.trust_region_radius);
    decrease_factor_ = round_real(sp_.radius_decrease_factor);
    bind(params);
    if (d_.gather) build_adjacency(true);
    // solver vectors start from zero so that excluded unknowns stay zero everywhere
    CD(cudaMemsetAsync(vec_block_, 0, vec_stride_ * 10, stream()));
    CD(cudaMemsetAsync(vecs_[V_P2], 0, vec_stride_, stream()));
    run_precompute();                        // gauss_newton.t:1190-1191
    prev_cost_ = compute_cost();
    {   // initX = X (copyUnknownwise)
        int dir = 0;
        void* args[] = {params_buf_.data(), &vecs_[V_INITX], &dir};
        launch_flat(fn("th_copy_x"), args);
    }
    log("Initial cost: %g\nInitial cost x 2: %g\n", prev_cost_, 2.0 * prev_cost_);
}

// finalize, gauss_newton.t:1200-1212
void Plan::finalize() {
    prev_cost_ = compute_cost();
    log("final cost=%g\n", prev_cost_);
    CD(cudaStreamSynchronize(stream()));
    if (cur_total_.a) span_end(cur_total_, ev_total_);
    CD(cudaStreamSynchronize(stream()));
    evaluate_timers();
    finalized_ = true;
}

double Plan::cost() {   // gauss_newton.t:1787-1793
    Scope scope(this);
    if (!finalized_) prev_cost_ = compute_cost();
    return prev_cost_;
}

void Plan::launch_tiled(int mode) {
    void* a[] = {params_buf_.data(), vecs_buf_.data(), maps_base(maps_buf_), &d_scalars_, &d_partials_, &mode, &peers_};
    const int v = use_tma_ ? 1 : 0;
    dim3 grid(tiled_grid_[v], 1, 1), block((unsigned)d_.tile[0], (unsigned)d_.tile[1], (unsigned)d_.tile[2]);
    launch(pcg_a_, grid, block, a, tiled_smem_[v]);
}

void Plan::linear_iteration(int l) {
    void* P = params_buf_.data();
    void* V = vecs_buf_.data();
    int zero = 0, one = 1;
    float qf = sp_.q_tolerance;
    double qd = sp_.q_tolerance;
    void* qtol = d_.is_double ? (void*)&qd : (void*)&qf;
    const bool reset = d_.lm && ((l + 1) % sp_.residual_reset_period) == 0;   // gauss_newton.t:1653-1660
    // ---- operator: Ap = (JtJ [+CtC]) p and alphaDenominator
    const bool nccl_scalars = d_.multi && !fused_;      // comparison path: PCG scalars over NCCL, separate push / close kernels
    if (d_.tiled) {
        launch_tiled(0);          // also forms p = z + beta p (PCGStep3 of the previous iteration)
        if (nccl_scalars) allreduce(offsetof(HScalars, aD), 1);
    } else if (d_.at_output) {
        void* a[] = {P, V, &d_scalars_, &d_partials_, &zero};
        launch_uw(fn("th_step1_uw"), a);
    } else if (d_.gather) {
        launch_gather(0);
        if (nccl_scalars) allreduce(offsetof(HScalars, aD), 1);
    } else {
        clear(vecs_[V_AP]);
        for (size_t g = 0; g < d_.groups.size(); ++g) {
            if (d_.groups[g].materialize) continue;   // (the front end lowers materialised J / J p groups to the gather schedule only)
            void* a[] = {P, V, &d_scalars_, &zero};
            launch_group(fn("th_applyjtj_g" + std::to_string(g)), (int)g, a);
        }
        void* a[] = {P, V, &d_scalars_, &d_partials_, &zero};
        launch_flat(fn("th_step1_finish"), a);
    }
    // ---- update of delta, r, z and the two dot products
    if (reset) {
        {
            void* a[] = {V, &d_scalars_};
            launch_flat(fn("th_step2_first"), a);
        }
        int add_ctc = 0;
        if (d_.tiled) {
            launch_tiled(1);      // (multi-GPU: the ghost layers of delta are maintained locally, th_pcg_b)
        } else if (d_.at_output) {
            void* a[] = {P, V, &d_scalars_, &d_partials_, &one};
            launch_uw(fn("th_step1_uw"), a);
        } else if (d_.gather) {
            launch_gather(1);        // materialised groups contribute nothing to A*delta (gauss_newton.t:1058-1065: no applyJTJ exists for them)
        } else {
            clear(vecs_[V_ADELTA]);
            for (size_t g = 0; g < d_.groups.size(); ++g) {
                if (d_.groups[g].materialize) continue;
                void* a[] = {P, V, &d_scalars_, &one};
                launch_group(fn("th_applyjtj_g" + std::to_string(g)), (int)g, a);
            }
            void* a[] = {P, V, &d_scalars_, &d_partials_, &one};
            launch_flat(fn("th_step1_finish"), a);
            add_ctc = 1;
        }
        void* a[] = {V, &d_scalars_, &d_partials_, &add_ctc, qtol, &d_flags_, &peers_, push_iter_.data()};
        launch_flat(fn("th_step2_second"), a);
    } else {
        void* a[] = {V, &d_scalars_, &d_partials_, qtol, &d_flags_, &peers_, push_iter_.data()};
        launch_flat(fn("th_pcg_b"), a);
    }
    if (d_.multi && fused_ && push_kernel_) {     // boundary layers of z -> the neighbours, in-kernel all-reduce of <z,r> and q, close
        void* a[] = {V, &d_scalars_, qtol, &d_flags_, &peers_, push_iter_.data()};
        launch(fn("th_push_close"), dim3(push_grid_), dim3(256), a);
    }
    if (nccl_scalars) {           // z ghost layers, global <z,r> and q, then close the iteration
        halo_push(V_Z, 1);
        allreduce(offsetof(HScalars, red), 2);
        void* a[] = {&d_scalars_, qtol, &d_flags_};
        launch(fn("th_mg_close"), dim3(1), dim3(1), a);
    }
    // ---- p = z + beta p (fused into the next th_pcg_a in the tiled schedule)
    if (!d_.tiled) {
        void* a[] = {V, &d_scalars_, qtol, &d_flags_};
        launch_flat(fn("th_step3"), a);
    }
}

// step, gauss_newton.t:1545-1785
int Plan::step(void** params) {
    Scope scope(this);
    bind(params);
    if (sp_.nIter >= sp_.nIterations) { finalize(); return 0; }
    void* P = params_buf_.data();
    void* V = vecs_buf_.data();
    span_begin(cur_iter_);
    span_begin(cur_phase_);
    ++epoch_;
    peers_.epoch = epoch_;
    const bool nccl_scalars = d_.multi && !fused_;
    int first = sp_.nIter == 0;
    if (d_.ncoef > 0) {   // PCG-invariant parts of J, once per nonlinear iteration
        void* a[] = {P};
        launch_uw(fn("th_precompute_coef"), a);
    }
    for (auto& c : d_.scoefs) {   // gather schedule: per-element invariants of each index space
        void* a[] = {P};
        launch(fn("th_precompute_scoef_s" + std::to_string(c.space)), dim3((unsigned)((d_.spaces[c.space].elements + 255) / 256)), dim3(256), a);
    }
    if (d_.at_output) {
        void* a[] = {P, V, &d_scalars_, &d_partials_, &first, &peers_, push_init_.data()};
        launch_uw(fn("th_init_uw"), a);
        if (nccl_scalars) {                   // z (= p0) ghost layers, global <r,p>
            halo_push(V_Z, 0);
            allreduce(offsetof(HScalars, rz), 1);
        }
    } else {
        if (d_.gather && gather_jtf_) {       // gathered over the adjacency lists: no clears, no atomics
            for (size_t s = 0; s < d_.spaces.size(); ++s) {
                void* a[] = {P, V, gather_buf_.data()};
                const long long threads = d_.spaces[s].elements * d_.spaces[s].lanes;
                launch(fn("th_gatherjtf_s" + std::to_string(s)), dim3((unsigned)((threads + 255) / 256)), dim3(256), a);
            }
        } else {
            clear(vecs_[V_R]);
            clear(vecs_[V_PRE]);
            for (size_t g = 0; g < d_.groups.size(); ++g) {
                void* a[] = {P, V};
                launch_group(fn("th_evaljtf_g" + std::to_string(g)), (int)g, a);
            }
        }
        for (auto& r : d_.replicated) {       // replicated unknowns: J^T F and diag(J^T J) summed over the ranks' residuals
            allreduce_vec(V_R, r.first, r.second);
            allreduce_vec(V_PRE, r.first, r.second);
        }
        void* a[] = {P, V, &d_scalars_, &d_partials_, &first, &peers_, push_init_.data()};
        launch_flat(fn("th_init_finish"), a);
        if (nccl_scalars) {                   // p0 at the ghost vertices, global <r,p>
            halo_push(V_P, 0);
            allreduce(offsetof(HScalars, rz), 1);
        }
        if (d_.gather)          // precomputeJ (gauss_newton.t:1019-1025, cusparseOuter :1332): store the partial derivatives
            for (size_t g = 0; g < d_.groups.size(); ++g) {
                if (d_.groups[g].materialize != 1) continue;
                void* aj[] = {P, gather_buf_.data()};
                launch_group(fn("th_computejv_g" + std::to_string(g)), (int)g, aj);
            }
    }
    span_end(cur_phase_, ev_setup_);
    span_begin(cur_phase_);
    // LM: the device decides when the linear solve ends (zeta test) and reports it through mapped
    // pinned memory; the host runs at most `depth` iterations ahead.  Before issuing iteration l it
    // waits until iteration l - depth + 1 has been closed and stops iff the exit was taken at or
    // before that iteration -- a function of device data only, so every rank of a multi-GPU solve
    // issues exactly the same sequence of kernels and collectives.
    const int depth = 4;
    if (use_graphs()) {
        // chunks of graph_chunk_ iterations, one graph launch each; at most two chunks in flight.  After the device's LM
        // exit the remaining kernels of a chunk return at once, and the host stops launching chunks as soon as it sees
        // the exit word (the fused multi-GPU kernels need no matching launch counts on the ranks).
        const int G = graph_chunk_;
        for (int l0 = 0; l0 < sp_.lIterations; l0 += G) {
            const int need = l0 - G;
            if (d_.lm && need >= 1) {
                bool stop = false;
                for (unsigned spins = 0;; ++spins) {
                    const long long pr = h_flags_->progress;
                    const int done_iters = (int)(pr >> 32) == epoch_ ? (int)(pr & 0xffffffff) : 0;
                    std::atomic_thread_fence(std::memory_order_acquire);
                    const long long ex = h_flags_->exit_word;
                    if ((int)(ex >> 32) == epoch_) { stop = true; break; }
                    if (done_iters >= need) break;
                    if ((spins & 1023) == 1023) {
                        const cudaError_t e = cudaStreamQuery(stream());
                        if (e != cudaSuccess && e != cudaErrorNotReady) fatal_cuda(e, "PCG iteration (asynchronous kernel failure)");
                    }
                    sched_yield();
                }
                if (stop) break;
            }
            run_chunk(l0, std::min(G, sp_.lIterations - l0));
        }
    } else
    for (int l = 0; l < sp_.lIterations; ++l) {
        const int need = l - depth + 1;
        if (d_.lm && need >= 1) {
            bool stop = false;
            for (unsigned spins = 0;; ++spins) {
                // the device publishes the exit word before the progress word (each with one 64-bit store)
                const long long pr = h_flags_->progress;
                const int done_iters = (int)(pr >> 32) == epoch_ ? (int)(pr & 0xffffffff) : 0;
                std::atomic_thread_fence(std::memory_order_acquire);
                const long long ex = h_flags_->exit_word;
                if ((int)(ex >> 32) == epoch_ && (int)(ex & 0xffffffff) <= need) { stop = true; break; }
                if (done_iters >= need) break;
                if ((spins & 1023) == 1023) {          // a faulted kernel never reports progress: surface the error
                    const cudaError_t e = cudaStreamQuery(stream());
                    if (e != cudaSuccess && e != cudaErrorNotReady) fatal_cuda(e, "PCG iteration (asynchronous kernel failure)");
                }
                sched_yield();
            }
            if (stop) break;
        }
        cur_tag_ = l;
        linear_iteration(l);
    }
    cur_tag_ = -1;
    span_end(cur_phase_, ev_linear_);
    span_begin(cur_phase_);
    int zero = 0;
    // (multi-GPU: the ghost layers of delta that the model cost and the update read are maintained locally, th_pcg_b)
    if (d_.lm) {   // computeModelCostChange + savePreviousUnknowns, gauss_newton.t:1694-1697
        if (d_.at_output) {
            void* a[] = {P, V, &d_scalars_, &d_partials_};
            launch_uw(fn("th_modelcost_uw"), a);
        } else
            for (size_t g = 0; g < d_.groups.size(); ++g) {
                int f = g == 0;
                void* a[] = {P, V, &d_scalars_, &d_partials_, &f};
                launch_group(fn("th_modelcost_g" + std::to_string(g)), (int)g, a);
            }
        if (d_.multi) allreduce(offsetof(HScalars, modelcost), 1);
        void* a[] = {P, &vecs_[V_PREVX], &zero};
        launch_flat(fn("th_copy_x"), a);
    }
    {
        void* a[] = {P, V};
        launch_flat(fn("th_update"), a);   // PCGLinearUpdate
    }
    run_precompute();                        // gauss_newton.t:1701
    int ret = 1;
    if (d_.lm) {
        const double new_cost = compute_cost();   // also brings modelcost and lin_done back
        last_linear_iterations = h_scalars_->lin_done;
        total_linear_iterations += (unsigned long long)h_scalars_->lin_done;
        epoch_lin_done_[epoch_] = h_scalars_->lin_done;
        const double model_cost = round_real(h_scalars_->modelcost);
        const double model_cost_change = round_real(prev_cost_ - model_cost);
        const double cost_change = round_real(prev_cost_ - new_cost);
        const double relative_decrease = round_real(cost_change / model_cost_change);
        log(" cost=%g model_cost=%g new cost=%g lin=%d rel=%g\n", prev_cost_, model_cost, new_cost, last_linear_iterations, relative_decrease);
        if (cost_change >= 0 && relative_decrease > round_real(sp_.min_relative_decrease)) {
            const double abs_tol = round_real(prev_cost_ * round_real(sp_.function_tolerance));
            if (cost_change <= abs_tol) {
                log("\nFunction tolerance reached (%g < %g), exiting\n", cost_change, abs_tol);
                span_end(cur_phase_, ev_finish_);
                span_end(cur_iter_, ev_iter_);
                finalize();
                return 0;
            }
            const double tmp_factor = 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3.0);
            radius_ = round_real(radius_ / std::fmax(1.0 / 3.0, tmp_factor));
            radius_ = round_real(std::fmin(radius_, (double)round_real(sp_.max_trust_region_radius)));
            decrease_factor_ = 2.0;
            prev_cost_ = new_cost;
        } else {
            int one = 1;
            void* a[] = {P, &vecs_[V_PREVX], &one};
            launch_flat(fn("th_copy_x"), a);   // revertUpdate
            run_precompute();                // gauss_newton.t:1748
            radius_ = round_real(radius_ / decrease_factor_);
            decrease_factor_ = round_real(2.0 * decrease_factor_);
            if (radius_ < round_real(sp_.min_trust_region_radius)) {
                log("\nTrust_region_radius is less than the min (%g), exiting\n", radius_);
                sp_.trust_region_radius = 10e4f;   // gauss_newton.t:1741
                span_end(cur_phase_, ev_finish_);
                span_end(cur_iter_, ev_iter_);
                finalize();
                return 0;
            }
            log("REVERT\n");
        }
        sp_.trust_region_radius = (float)radius_;
        log(" trust_region_radius=%g\n", radius_);
    } else {
        last_linear_iterations = sp_.lIterations;
        total_linear_iterations += (unsigned long long)sp_.lIterations;
        epoch_lin_done_[epoch_] = sp_.lIterations;
    }
    sp_.nIter += 1;
    span_end(cur_phase_, ev_finish_);
    span_end(cur_iter_, ev_iter_);
    if (sp_.max_solver_time_in_seconds > 0.0f) {
        CD(cudaStreamSynchronize(stream()));
        const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start_).count();
        if (el > sp_.max_solver_time_in_seconds) {
            log("\nTime exceeded max (%g > %g), exiting\n", el, (double)sp_.max_solver_time_in_seconds);
            finalize();
            return 0;
        }
    }
    return ret;
}

void Plan::solve(void** params) {   // thallo.t:5980-5983
    Scope scope(this);
    init(params);
    while (step(params)) {}
}

// setSolverParameter / getSolverParameter, gauss_newton.t:1828-1862
#define TH_PARAMS(X) \
    X(min_relative_decrease, float) X(min_trust_region_radius, float) X(max_trust_region_radius, float) \
    X(q_tolerance, float) X(function_tolerance, float) X(trust_region_radius, float) X(radius_decrease_factor, float) \
    X(min_lm_diagonal, float) X(max_lm_diagonal, float) X(max_solver_time_in_seconds, float) \
    X(residual_reset_period, int) X(nIter, int) X(nIterations, int) X(lIterations, int)

void Plan::set_parameter(const char* name, const void* value) {
#define X(f, T) if (strcmp(#f, name) == 0) { sp_.f = *(const T*)value; return; }
    TH_PARAMS(X)
#undef X
    log("Warning: tried to set nonexistent solver parameter %s\n", name);
}
void Plan::get_parameter(const char* name, void* value) {
#define X(f, T) if (strcmp(#f, name) == 0) { *(T*)value = sp_.f; return; }
    TH_PARAMS(X)
#undef X
    log("Warning: tried to get nonexistent solver parameter %s\n", name);
}

long long Plan::read_vector(const char* name, void* dst, long long count) {
    Scope scope(this);
    for (int i = 0; i < kNumVecs; ++i) {
        if (strcmp(name, kVecNames[i]) == 0) {
            const long long n = std::min<long long>(count, d_.nunk);
            CD(cudaMemcpyAsync(dst, vecs_[i], (size_t)n * real_size_, cudaMemcpyDeviceToHost, stream()));
            CD(cudaStreamSynchronize(stream()));
            return n;
        }
    }
    return 0;
}

void* Plan::vector_pointer(const char* name) {
    for (int i = 0; i < kNumVecs; ++i)
        if (strcmp(name, kVecNames[i]) == 0) return vecs_[i];
    return nullptr;
}

// The Jacobian of one residual group at the current unknowns, as the reference materialises it (generateDumpJ,
// gauss_newton.t:325-487): per residual element its rows term by term, row k holding row_nnz[k] (value, column)
// pairs; column = flat index of the unknown scalar (imageOffset + channels*idx + ch) or -1 outside the domain.
long long Plan::export_jacobian(int g, void* host_vals, long long* host_cols, long long capacity) {
    if (g < 0 || g >= (int)d_.groups.size() || params_buf_.empty()) return -1;
    const long long n = d_.groups[g].count * (long long)d_.groups[g].nnz;
    if (n > capacity) return -1;
    if (n == 0) return 0;
    Scope scope(this);
    void* dv = nullptr; long long* dc = nullptr;
    CD(cudaMalloc(&dv, (size_t)n * real_size_));
    CD(cudaMalloc((void**)&dc, (size_t)n * sizeof(long long)));
    void* a[] = {params_buf_.data(), &dv, &dc};
    launch_group(fn("th_computej_g" + std::to_string(g)), g, a);
    CD(cudaMemcpyAsync(host_vals, dv, (size_t)n * real_size_, cudaMemcpyDeviceToHost, stream()));
    CD(cudaMemcpyAsync(host_cols, dc, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, stream()));
    CD(cudaStreamSynchronize(stream()));
    cudaFree(dv); cudaFree(dc);
    return n;
}

}  // namespace thallo
