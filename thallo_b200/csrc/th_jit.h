// NVRTC compile + module load through the CUDA driver API (entry points fetched from the
// runtime, so the library has no link-time dependency on libcuda and loads on a CPU box).
// Replaces reference API/src/util.t:868 / cuda_util.t:470 (terralib.cudacompile) and the
// PTX -> cuModuleLoad step the Terra runtime performs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace thallo {

struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool ok = false;
    static const DriverApi& get();   // initialises the runtime (primary context) on first use
};

// Compile `source` for sm_100a.  include_dir holds thallo_prelude.cuh / thallo_kernels.cuh.
// Returns true and fills cubin; log always receives the compiler log.
bool compile_cubin(const std::string& source, const std::string& include_dir, std::vector<char>& cubin, std::string& log,
                   const std::vector<std::string>& extra_opts = {});

// Directory of the skeleton headers: $THALLO_B200_SKELETON_DIR or <dir of this .so>/skeleton.
std::string skeleton_dir();
std::string library_dir();

}  // namespace thallo
