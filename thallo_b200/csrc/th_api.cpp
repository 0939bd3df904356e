// C ABI of thallo_b200: the twelve Thallo.h entry points (reference
// API/release/include/Thallo.h:41-106, forwarders createwrapper.t:226-232) plus the
// extension seam declared in include/thallo_b200.h.
#include <fcntl.h>
#include <spawn.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>
#include <map>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/thallo_b200.h"
#include "th_plan.h"

using namespace thallo;

struct Thallo_Problem {
    std::string filename, kind;
    bool from_source = false;
    std::string descriptor, source;
    bool deleted = false;
};
struct Thallo_State {
    StateOptions opts;
    std::vector<Thallo_Problem*> problems;
};
struct Thallo_Plan {
    Plan* plan;
};

static thread_local std::string g_last_error;
static void set_error(const std::string& s) {
    g_last_error = s;
    fprintf(stderr, "thallo_b200: %s\n", s.c_str());
}

static std::string read_all(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// Runs the energy front end as a child process:  <python> -m thallo_b200.frontend --dims-on-stdin ...
// (the reference evaluates the .t file inside its embedded Lua VM at ProblemPlan time,
// thallo.t:1359-1373,1384-1434; we have no VM in the library, so the front end is a tool).
// ONE process per plan, started with posix_spawnp and an argument vector (no shell: a quote or a space in the caller's
// path is just a character); the child reports how many dimensions the energy declares, the library answers with that
// many entries of the caller's array, the child lowers.  Results are cached in-process by (path, mtime, size, dims,
// kind, precision, mode), so re-planning the same energy -- the reference's tests/create_delete_cycle -- starts no
// interpreter at all.
struct LoweredEnergy { std::string desc, src; };
static std::map<std::string, LoweredEnergy>& lowering_cache() {
    static std::map<std::string, LoweredEnergy> c;
    return c;
}
static std::map<std::string, int>& ndims_cache() {
    static std::map<std::string, int> c;
    return c;
}
static std::string file_stamp(const std::string& path) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return path + "|absent";
    std::ostringstream o;
    o << path << "|" << (long long)st.st_mtime << "|" << (long long)st.st_size;
    return o.str();
}
static void rm_rf_dir(const std::string& dir) {       // the temporary directory holds plain files only
    for (const char* f : {"/plan.desc", "/energy.cu", "/log.txt", "/ndims.txt"}) unlink((dir + f).c_str());
    rmdir(dir.c_str());
}
extern char** environ;
static bool run_frontend(const Thallo_Problem* pr, const StateOptions& o, const unsigned int* dims, std::string& desc, std::string& src) {
    const bool as_committed = getenv("THALLO_LM_AS_COMMITTED") != nullptr;
    const std::string stamp = file_stamp(pr->filename) + "|" + pr->kind + "|" + (o.init.doublePrecision ? "d" : "f") + (as_committed ? "|gn" : "");
    auto nd_it = ndims_cache().find(stamp);
    if (nd_it != ndims_cache().end()) {
        std::ostringstream key;
        key << stamp;
        for (int i = 0; i < nd_it->second; ++i) key << "|" << dims[i];
        auto hit = lowering_cache().find(key.str());
        if (hit != lowering_cache().end()) { desc = hit->second.desc; src = hit->second.src; return true; }
    }
    const char* py = getenv("THALLO_B200_PYTHON");
    const std::string root = getenv("THALLO_B200_ROOT") ? getenv("THALLO_B200_ROOT") : library_dir() + "/../..";
    char tmpl[] = "/tmp/thallo_b200_XXXXXX";
    if (!mkdtemp(tmpl)) { set_error("mkdtemp failed"); return false; }
    const std::string out(tmpl), logf = out + "/log.txt";
    int to_child[2], from_child[2];
    if (pipe(to_child) != 0 || pipe(from_child) != 0) { set_error("pipe failed"); rm_rf_dir(out); return false; }
    std::vector<std::string> args = {py ? py : "python3", "-m", "thallo_b200.frontend", "--energy", pr->filename, "--kind", pr->kind,
                                     "--double", o.init.doublePrecision ? "1" : "0", "--out", out, "--dims-on-stdin"};
    if (as_committed) args.push_back("--lm-as-committed");
    std::vector<char*> argv;
    for (auto& a : args) argv.push_back(const_cast<char*>(a.c_str()));
    argv.push_back(nullptr);
    // environment: the caller's, with the package root in front of PYTHONPATH
    std::vector<std::string> envs;
    std::string pp = "PYTHONPATH=" + root;
    for (char** e = environ; e && *e; ++e) {
        if (strncmp(*e, "PYTHONPATH=", 11) == 0) { if ((*e)[11]) pp += std::string(":") + (*e + 11); }
        else envs.push_back(*e);
    }
    envs.push_back(pp);
    std::vector<char*> envp;
    for (auto& e : envs) envp.push_back(const_cast<char*>(e.c_str()));
    envp.push_back(nullptr);
    posix_spawn_file_actions_t fa;
    posix_spawn_file_actions_init(&fa);
    posix_spawn_file_actions_adddup2(&fa, to_child[0], 0);
    posix_spawn_file_actions_adddup2(&fa, from_child[1], 1);
    posix_spawn_file_actions_addopen(&fa, 2, logf.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);
    for (int fd : {to_child[0], to_child[1], from_child[0], from_child[1]}) posix_spawn_file_actions_addclose(&fa, fd);
    pid_t pid = 0;
    const int rc = posix_spawnp(&pid, argv[0], &fa, nullptr, argv.data(), envp.data());
    posix_spawn_file_actions_destroy(&fa);
    close(to_child[0]);
    close(from_child[1]);
    bool ok = false;
    int nd = 0;
    if (rc != 0) {
        set_error(std::string("could not start the energy front end (") + argv[0] + "): " + strerror(rc));
        close(to_child[1]); close(from_child[0]);
        rm_rf_dir(out);
        return false;
    }
    {   // "ndims N\n" from the child (EOF = it failed before getting there)
        std::string line;
        char c;
        while (read(from_child[0], &c, 1) == 1 && c != '\n') line.push_back(c);
        if (line.compare(0, 6, "ndims ") == 0) nd = atoi(line.c_str() + 6);
    }
    if (nd > 0) {
        std::ostringstream d;
        for (int i = 0; i < nd; ++i) d << (i ? "," : "") << dims[i];
        d << "\n";
        const std::string ds = d.str();
        if (write(to_child[1], ds.data(), ds.size()) != (ssize_t)ds.size()) nd = 0;
    }
    close(to_child[1]);
    close(from_child[0]);
    int status = 0;
    while (waitpid(pid, &status, 0) < 0 && errno == EINTR) {}
    if (nd <= 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        set_error("energy front end failed for '" + pr->filename + "':\n" + read_all(logf));
    } else {
        desc = read_all(out + "/plan.desc");
        src = read_all(out + "/energy.cu");
        ok = !desc.empty() && !src.empty();
        if (!ok) set_error("energy front end produced no output for '" + pr->filename + "'");
    }
    rm_rf_dir(out);
    if (ok) {
        ndims_cache()[stamp] = nd;
        std::ostringstream key;
        key << stamp;
        for (int i = 0; i < nd; ++i) key << "|" << dims[i];
        lowering_cache()[key.str()] = LoweredEnergy{desc, src};
    }
    return ok;
}

extern "C" {

Thallo_State* Thallo_NewState(Thallo_InitializationParameters params) {
    Thallo_State* s = new Thallo_State();
    s->opts.init = params;
    if (params.cpuOnly) set_error("cpuOnly=1 requested: thallo_b200 has no CPU backend; plans will run on the GPU");
    return s;
}

Thallo_Problem* Thallo_ProblemDefine(Thallo_State* state, const char* filename, const char* solverkind) {
    if (!state || !filename || !solverkind) return nullptr;
    if (strcmp(solverkind, "gauss_newton") != 0 && strcmp(solverkind, "levenberg_marquardt") != 0) {
        set_error(std::string("unknown solver kind '") + solverkind + "' (thallo.t:74)");
        return nullptr;
    }
    Thallo_Problem* p = new Thallo_Problem();
    p->filename = filename;
    p->kind = solverkind;
    state->problems.push_back(p);
    return p;
}

Thallo_Problem* ThalloB200_ProblemDefineFromSource(Thallo_State* state, const char* descriptor, const char* cuda_source,
                                                   const char* solverkind) {
    if (!state || !descriptor || !cuda_source || !solverkind) return nullptr;
    Thallo_Problem* p = new Thallo_Problem();
    p->from_source = true;
    p->descriptor = descriptor;
    p->source = cuda_source;
    p->kind = solverkind;
    state->problems.push_back(p);
    return p;
}

void Thallo_ProblemDelete(Thallo_State* state, Thallo_Problem* problem) {
    (void)state;
    if (problem) problem->deleted = true;   // tombstone only, like thallo.t:5954-5961
}

Thallo_Plan* Thallo_ProblemPlan(Thallo_State* state, Thallo_Problem* problem, unsigned int* dimensions) {
    if (!state || !problem || problem->deleted) return nullptr;
    std::string desc_text, src;
    if (problem->from_source) {
        desc_text = problem->descriptor;
        src = problem->source;
    } else {
        if (!dimensions) { set_error("Thallo_ProblemPlan: dimensions is NULL"); return nullptr; }
        if (!run_frontend(problem, state->opts, dimensions, desc_text, src)) return nullptr;
    }
    PlanDesc d;
    std::string err;
    if (!parse_descriptor(desc_text, d, err)) { set_error(err); return nullptr; }
    if (d.lm && state->opts.init.verbosityLevel > 0) {
        static bool told = false;
        if (!told) {
            told = true;
            fprintf(stderr, "thallo_b200: note: \"levenberg_marquardt\" runs Levenberg-Marquardt as written in gauss_newton.t; the reference snapshot "
                            "runs Gauss-Newton for this kind string (thallo.t:463). THALLO_LM_AS_COMMITTED=1 reproduces the snapshot.\n");
        }
    }
    if (problem->from_source && dimensions) {
        for (size_t i = 0; i < d.dims.size(); ++i)
            if ((long long)dimensions[i] != d.dims[i]) {
                set_error("dimensions passed to Thallo_ProblemPlan differ from the ones the energy was lowered for");
                return nullptr;
            }
    }
    if ((d.is_double ? 1 : 0) != (state->opts.init.doublePrecision ? 1 : 0)) {
        set_error("energy was lowered for a different precision than the state's doublePrecision");
        return nullptr;
    }
    Plan* plan = new Plan(&state->opts, d, src);
    if (!plan->ok()) {
        set_error(plan->error());
        delete plan;
        return nullptr;
    }
    Thallo_Plan* h = new Thallo_Plan();
    h->plan = plan;
    return h;
}

void Thallo_PlanFree(Thallo_State* state, Thallo_Plan* plan) {
    (void)state;
    if (!plan) return;
    delete plan->plan;
    delete plan;
}

void Thallo_SetSolverParameter(Thallo_State*, Thallo_Plan* plan, const char* name, void* value) {
    if (plan && name && value) plan->plan->set_parameter(name, value);
}
void Thallo_GetSolverParameter(Thallo_State*, Thallo_Plan* plan, const char* name, void* value) {
    if (plan && name && value) plan->plan->get_parameter(name, value);
}
void Thallo_ProblemSolve(Thallo_State*, Thallo_Plan* plan, void** problemparams) { plan->plan->solve(problemparams); }
void Thallo_ProblemInit(Thallo_State*, Thallo_Plan* plan, void** problemparams) { plan->plan->init(problemparams); }
int Thallo_ProblemStep(Thallo_State*, Thallo_Plan* plan, void** problemparams) { return plan->plan->step(problemparams); }
double Thallo_ProblemCurrentCost(Thallo_State*, Thallo_Plan* plan) { return plan->plan->cost(); }
void Thallo_GetPerformanceSummary(Thallo_State*, Thallo_Plan* plan, Thallo_PerformanceSummary* summary) {
    if (plan && summary) plan->plan->summary(summary);
}

int ThalloB200_CompileOnly(const char* cuda_source, char* log, unsigned long log_capacity, unsigned long* cubin_size) {
    std::vector<char> cubin;
    std::string clog;
    const bool ok = compile_cubin(cuda_source ? cuda_source : "", skeleton_dir(), cubin, clog);
    if (log && log_capacity) {
        strncpy(log, clog.c_str(), log_capacity - 1);
        log[log_capacity - 1] = 0;
    }
    if (cubin_size) *cubin_size = (unsigned long)cubin.size();
    return ok ? 0 : 1;
}

void ThalloB200_SetStream(Thallo_State* state, void* cuda_stream) {
    if (state) state->opts.stream = (cudaStream_t)cuda_stream;
}
unsigned long long ThalloB200_PlanLaunchCount(Thallo_State*, Thallo_Plan* plan) { return plan ? plan->plan->launches : 0; }
int ThalloB200_PlanLastLinearIterations(Thallo_State*, Thallo_Plan* plan) { return plan ? plan->plan->last_linear_iterations : 0; }
unsigned long long ThalloB200_PlanTotalLinearIterations(Thallo_State*, Thallo_Plan* plan) {
    return plan ? plan->plan->total_linear_iterations : 0;
}
long long ThalloB200_PlanExportJacobian(Thallo_State*, Thallo_Plan* plan, int group, void* host_vals, long long* host_cols,
                                        long long capacity) {
    return plan && host_vals && host_cols ? plan->plan->export_jacobian(group, host_vals, host_cols, capacity) : -1;
}
long long ThalloB200_PlanReadVector(Thallo_State*, Thallo_Plan* plan, const char* name, void* host_dst, long long count) {
    return plan ? plan->plan->read_vector(name, host_dst, count) : 0;
}
long long ThalloB200_PlanKernelTimes(Thallo_State*, Thallo_Plan* plan, char* buf, long long capacity) {
    if (!plan || !buf || capacity <= 0) return 0;
    const std::string t = plan->plan->kernel_times();
    const long long n = std::min<long long>((long long)t.size(), capacity - 1);
    memcpy(buf, t.data(), (size_t)n);
    buf[n] = 0;
    return n;
}
int ThalloB200_NcclUniqueId(void* id, int capacity) { return thallo::nccl_unique_id(id, capacity); }
int ThalloB200_PlanInitComm(Thallo_State*, Thallo_Plan* plan, const void* nccl_id, int rank, int world) {
    if (!plan || !nccl_id) return 1;
    const int r = plan->plan->comm_init(nccl_id, rank, world);
    if (r) set_error(plan->plan->error());
    return r;
}
int ThalloB200_PlanIpcHandle(Thallo_State*, Thallo_Plan* plan, void* handle64, long long* slow_extent) {
    return plan && handle64 ? plan->plan->ipc_handle(handle64, slow_extent) : 1;
}
int ThalloB200_PlanConnect(Thallo_State*, Thallo_Plan* plan, const void* handle_lo, long long extent_lo, const void* handle_hi,
                           long long extent_hi) {
    return plan ? plan->plan->connect(handle_lo, extent_lo, handle_hi, extent_hi) : 1;
}
int ThalloB200_PlanConnectGraph(Thallo_State*, Thallo_Plan* plan, const void* handle_lo, long long extent_lo, long long width_lo,
                                const void* handle_hi, long long extent_hi, long long width_hi) {
    return plan ? plan->plan->connect_graph(handle_lo, extent_lo, width_lo, handle_hi, extent_hi, width_hi) : 1;
}
void* ThalloB200_PlanVectorPointer(Thallo_State*, Thallo_Plan* plan, const char* name) {
    return plan && name ? plan->plan->vector_pointer(name) : nullptr;
}
int ThalloB200_PlanPeerInfo(Thallo_State*, Thallo_Plan* plan, long long* info4) {
    return plan && info4 ? plan->plan->peer_info(info4) : 1;
}
int ThalloB200_PlanConnectAll(Thallo_State*, Thallo_Plan* plan, int world, const void* handles64, const long long* infos4) {
    if (!plan || !handles64 || !infos4) return 1;
    const int rc = plan->plan->connect_all(world, handles64, infos4);
    if (rc) set_error(plan->plan->error());
    return rc;
}
int ThalloB200_WarpSelfTest(int which, int nkeys, double* out, int capacity) {
    if (!out || capacity <= 0 || which < 0 || which > 2) return -1;
    if (nkeys < 1 || nkeys > 32) nkeys = 4;
    const DriverApi& drv = DriverApi::get();
    if (!drv.ok) { set_error("no CUDA device / driver available"); return -1; }
    std::vector<char> cubin;
    std::string log;
    if (!compile_cubin("#define TH_WARP_KAT 1\n#include \"thallo_warp.cuh\"\n", skeleton_dir(), cubin, log)) {
        set_error("warp self-test: NVRTC failed:\n" + log);
        return -1;
    }
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    const char* names[3] = {"th_kat_ballot", "th_kat_get_peers", "th_kat_reduce_peers"};
    if (drv.ModuleLoadData(&mod, cubin.data()) != CUDA_SUCCESS || drv.ModuleGetFunction(&fn, mod, names[which]) != CUDA_SUCCESS) {
        set_error("warp self-test: module load failed");
        if (mod) drv.ModuleUnload(mod);
        return -1;
    }
    unsigned char* dbuf = nullptr;                    // [32 x float/unsigned | 32 x double]
    const size_t bytes = 32 * 4 + 32 * 8;
    int written = -1;
    if (cudaMalloc((void**)&dbuf, bytes) == cudaSuccess && cudaMemset(dbuf, 0, bytes) == cudaSuccess) {
        void* a0 = dbuf;
        void* a1 = dbuf + 32 * 4;
        void* args[3] = {&a0, &a1, &nkeys};
        if (drv.LaunchKernel(fn, 1, 1, 1, 32, 1, 1, 0, nullptr, args, nullptr) == CUDA_SUCCESS && cudaDeviceSynchronize() == cudaSuccess) {
            unsigned char h[32 * 4 + 32 * 8];
            if (cudaMemcpy(h, dbuf, bytes, cudaMemcpyDeviceToHost) == cudaSuccess) {
                written = 0;
                if (which < 2) out[written++] = (double)*(const unsigned*)h;
                else {
                    for (int k = 0; k < nkeys && written < capacity; ++k) out[written++] = (double)((const float*)h)[k];
                    for (int k = 0; k < nkeys && written < capacity; ++k) out[written++] = ((const double*)(h + 32 * 4))[k];
                }
            }
        } else set_error("warp self-test: launch failed");
    }
    if (dbuf) cudaFree(dbuf);
    drv.ModuleUnload(mod);
    return written;
}
const char* ThalloB200_LastError(void) { return g_last_error.c_str(); }
const char* ThalloB200_Version(void) { return "thallo_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
