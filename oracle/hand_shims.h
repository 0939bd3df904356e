// ORACLE (test infrastructure): host equivalents of the few device-only constructs the reference's hand-written solver
// headers touch (examples/*/src/*SolverUtil.h: warp shuffles, __syncthreads, dynamic shared memory, the built-in thread
// indices).  None of them is executed by the drivers, which call the per-variable equation functions directly.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

static inline float __shfl_down(float v, int, int) { return v; }
static inline void __syncthreads() {}
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float atomicAdd(float* a, float v) { float o = *a; *a += v; return o; }
struct ThUint3 { unsigned x, y, z; };
static ThUint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#undef __shared__
#define __shared__
float bucket[2048];

