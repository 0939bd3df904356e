// Stand-alone C++ caller written against include/Thallo.h only, shaped like the reference's
// test programs (reference tests/minimal/main.cpp:9-33,49-72): builds the MSVC-rand() target on
// the host, solves with Thallo_ProblemSolve via the *file name* of the reference energy, and
// writes the 8-bit result so that the Python test can compare it with the golden image.
// Build: g++ minimal_main.cpp -I../../include -L../../thallo_b200/lib -lThallo -lcudart
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>
extern "C" {
#include "Thallo.h"
}

static unsigned int msvc_state = 1;
static int msvc_rand() {
    msvc_state = msvc_state * 214013u + 2531011u;
    return (int)((msvc_state >> 16) & 0x7fff);
}

int main(int argc, char** argv) {
    const char* energy = argc > 1 ? argv[1] : "tests/minimal/laplacian.t";
    const char* outpath = argc > 2 ? argv[2] : "result.u8";
    const int dim = 512;
    std::vector<float> scratch(dim * dim);
    for (int i = 0; i < dim * dim; ++i) scratch[i] = (float)((double)msvc_rand() / 32767.0);
    float *target, *unknown;
    size_t fSize = dim * dim * sizeof(float);
    cudaMalloc((void**)&target, fSize);
    cudaMalloc((void**)&unknown, fSize);
    cudaMemcpy(target, scratch.data(), fSize, cudaMemcpyHostToDevice);
    cudaMemcpy(unknown, target, fSize, cudaMemcpyDeviceToDevice);

    Thallo_InitializationParameters param = {};
    param.verbosityLevel = 1;
    param.timingLevel = 2;
    Thallo_State* state = Thallo_NewState(param);
    Thallo_Problem* problem = Thallo_ProblemDefine(state, energy, "gauss_newton");
    unsigned int dims[] = {(unsigned)dim, (unsigned)dim};
    Thallo_Plan* plan = Thallo_ProblemPlan(state, problem, dims);
    if (!plan) { fprintf(stderr, "plan failed\n"); return 2; }
    void* problem_data[] = {unknown, target};
    Thallo_ProblemSolve(state, plan, problem_data);
    double cost = Thallo_ProblemCurrentCost(state, plan);
    Thallo_PerformanceSummary sum;
    Thallo_GetPerformanceSummary(state, plan, &sum);
    Thallo_PlanFree(state, plan);
    Thallo_ProblemDelete(state, problem);

    cudaMemcpy(scratch.data(), unknown, fSize, cudaMemcpyDeviceToHost);
    std::vector<unsigned char> out(dim * dim);
    for (int i = 0; i < dim * dim; ++i) out[i] = (unsigned char)(scratch[i] * 255);
    FILE* f = fopen(outpath, "wb");
    fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    printf("\nminimal %g  (total %.3f ms, %u nonlinear iterations)\n", cost, sum.total.meanMS, sum.nonlinearIteration.count);
    cudaFree(target);
    cudaFree(unknown);
    return 0;
}
