#include "th_jit.h"

#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <mutex>
#include <sstream>

namespace thallo {

static void* entry(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPointByVersion(name, &fn, 12000, cudaEnableDefault, &st) != cudaSuccess ||
        st != cudaDriverEntryPointSuccess)
        return nullptr;
    return fn;
}

const DriverApi& DriverApi::get() {
    static DriverApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        if (cudaFree(nullptr) != cudaSuccess) {   // creates / binds the primary context
            cudaGetLastError();
            return;
        }
#define TH_ENTRY(field, sym) api.field = reinterpret_cast<decltype(api.field)>(entry(sym))
        TH_ENTRY(ModuleLoadData, "cuModuleLoadData");
        TH_ENTRY(ModuleUnload, "cuModuleUnload");
        TH_ENTRY(ModuleGetFunction, "cuModuleGetFunction");
        TH_ENTRY(LaunchKernel, "cuLaunchKernel");
        TH_ENTRY(FuncSetAttribute, "cuFuncSetAttribute");
        TH_ENTRY(GetErrorString, "cuGetErrorString");
        TH_ENTRY(OccupancyMaxActiveBlocks, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
        TH_ENTRY(TensorMapEncodeTiled, "cuTensorMapEncodeTiled");
#undef TH_ENTRY
        api.ok = api.ModuleLoadData && api.ModuleGetFunction && api.LaunchKernel && api.ModuleUnload;
    });
    return api;
}

std::string library_dir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&library_dir), &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t s = p.find_last_of('/');
        return s == std::string::npos ? std::string(".") : p.substr(0, s);
    }
    return ".";
}

std::string skeleton_dir() {
    if (const char* e = getenv("THALLO_B200_SKELETON_DIR")) return e;
    const std::string in_tree = library_dir() + "/../csrc/skeleton";     // thallo_b200/lib -> thallo_b200/csrc/skeleton
    struct stat st;
    if (stat((in_tree + "/thallo_kernels.cuh").c_str(), &st) == 0) return in_tree;
    return library_dir() + "/skeleton";                                   // installed layout
}

static std::string cache_dir() {
    if (const char* e = getenv("THALLO_B200_CACHE_DIR")) return e;
    return library_dir() + "/.jit_cache";
}

static std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

bool compile_cubin(const std::string& source, const std::string& include_dir, std::vector<char>& cubin, std::string& log,
                   const std::vector<std::string>& extra_opts) {
    // cache key: source + skeleton headers + options
    std::string key_src = source + "\n//--\n" + read_file(include_dir + "/thallo_prelude.cuh") + "\n//--\n" +
                          read_file(include_dir + "/thallo_access.cuh") + "\n//--\n" + read_file(include_dir + "/thallo_kernels.cuh") +
                          "\n//--\n" + read_file(include_dir + "/thallo_warp.cuh");
    for (auto& o : extra_opts) key_src += "\n" + o;
    const size_t h1 = std::hash<std::string>{}(key_src);
    const size_t h2 = std::hash<std::string>{}(key_src + "#salt");
    char name[64];
    snprintf(name, sizeof name, "%016zx%016zx.cubin", h1, h2);
    const bool use_cache = !getenv("THALLO_B200_NO_CACHE");
    const std::string cdir = cache_dir(), cpath = cdir + "/" + name;
    if (use_cache) {
        std::ifstream f(cpath, std::ios::binary);
        if (f) {
            cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
            if (!cubin.empty()) {
                log = "(cubin cache hit " + cpath + ")";
                return true;
            }
        }
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, source.c_str(), "thallo_energy.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device",
                                     "-I" + include_dir};
    for (auto& o : extra_opts) opts.push_back(o);
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    nvrtcResult r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
    size_t lsz = 0;
    nvrtcGetProgramLogSize(prog, &lsz);
    log.resize(lsz);
    if (lsz) nvrtcGetProgramLog(prog, &log[0]);
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return false;
    }
    size_t csz = 0;
    if (nvrtcGetCUBINSize(prog, &csz) != NVRTC_SUCCESS || csz == 0) {
        log += "\nnvrtcGetCUBINSize failed";
        nvrtcDestroyProgram(&prog);
        return false;
    }
    cubin.resize(csz);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    if (use_cache) {
        mkdir(cdir.c_str(), 0755);
        std::string tmp = cpath + ".tmp" + std::to_string((long)getpid());
        std::ofstream f(tmp, std::ios::binary);
        if (f) {
            f.write(cubin.data(), (std::streamsize)cubin.size());
            f.close();
            rename(tmp.c_str(), cpath.c_str());
        }
    }
    return true;
}

}  // namespace thallo
