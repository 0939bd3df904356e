// thallo_b200 solver skeleton, part 1: types, math, plan-wide tables.
// Included by every generated per-energy translation unit *before* the generated
// device functions (namespace th).  Compiled by NVRTC (no host headers available)
// or nvcc for sm_100a.
//
// Reference roles replaced: API/src/util.t:152-199 (gpuMath), :203-301 (Vector),
// API/src/precision.t:3-7 (thallo_float), thallo.t:245-258 (ProblemParameters).
#pragma once

#if TH_DOUBLE
typedef double real;
#else
typedef float real;
#endif
#if TH_DOUBLE
typedef double4 real4;
#else
typedef float4 real4;
#endif
// read-only 4-scalar chunk c of a 16-byte aligned array
__device__ __forceinline__ real4 th_ld4(const real* p, int c) {
#if TH_DOUBLE
    const double2 a = __ldg(((const double2*)p) + 2 * c), b = __ldg(((const double2*)p) + 2 * c + 1);
    return make_double4(a.x, a.y, b.x, b.y);
#else
    return __ldg(((const float4*)p) + c);
#endif
}
typedef real th_real;
typedef float th_float;
typedef unsigned char th_uchar;
typedef int th_int;

#define TH_INF ((real)__int_as_float(0x7f800000))
#define TH_NAN ((real)__int_as_float(0x7fffffff))
#define TH_MAXD 3

// Problem parameters as the generated code sees them: device pointers by slot
// (images, unknowns, sparse index arrays) and scalar Params by slot; LM scalars
// appended (thallo.t:1584-1589).
struct Params {
    void* ptr[TH_NPTR];
    real sc[TH_NSC];
    real trust_region_radius;
    real radius_decrease_factor;
    real min_lm_diagonal;
    real max_lm_diagonal;
};

// Solver vectors (gauss_newton.t:282-309 PlanData): every one is a flat array of
// TH_NUNK reals, unknown images back to back, AoS per element.
struct Vecs {
    real* delta; real* r; real* b; real* Adelta; real* z; real* p; real* Ap;
    real* CtC; real* pre; real* SSq;
    real* p2;      // second search-direction buffer: the tiled operator kernel ping-pongs p (see th_pcg_a)
};

// Device-resident scalars.  rz[] double-buffers the CG numerator so that no
// device-to-device copy is needed between iterations (the reference copies
// scanBetaNumerator -> scanAlphaNumerator, gauss_newton.t:1665).
struct ThScalars {
    double rz[2];
    double aD;
    double q;
    double Q0;
    double cost;
    double modelcost;
    double spare;
    double red[2];        // multi-GPU: this rank's partial <z,r> and q, all-reduced before the iteration is closed
    unsigned int ticket[8];
    int it;
    int done;
    int lin_done;
    int epoch;            // number of the nonlinear step, set by the PCGInit kernel: the kernels of the PCG iteration take no
                          // per-step argument, so a captured CUDA graph of the iteration can be replayed in every step
};

// Host-visible progress flags (pinned, mapped): two 64-bit words, each (epoch << 32) | iteration, each
// published with a single store -- `progress` = the iteration just closed, `exit_word` = the iteration at
// which the LM early exit was taken (0 in the low half = not taken in this epoch).
struct ThHostFlags { long long progress; long long exit_word; };

// ---- multi-GPU peer memory (one process per GPU; every rank maps every other rank's solver-vector block with
// CUDA IPC).  The PCG scalars are all-reduced INSIDE the kernel that produces them: the last CTA of the grid
// reduction stores this rank's partial sums into a mailbox in every peer's memory over NVLink, then waits for
// the peers' values in its own mailbox and adds them up in rank order (bit-identical totals on every rank).
// Boundary layers of z / p are stored straight into the neighbours' ghost layers by the kernel that computes
// them (ThPush), ordered before the mailbox flag by a system-scope fence.
#define TH_MAXRANKS 16
struct ThMail { unsigned long long q[4]; };       // two (value bits, sequence number) pairs, each written with ONE 16-byte store
enum { TH_MAIL_A = 0, TH_MAIL_B = 1, TH_MAIL_INIT = 2, TH_MAIL_KINDS = 4 };
#define TH_MAIL_BYTES (TH_MAIL_KINDS * 2 * TH_MAXRANKS * (int)sizeof(ThMail))
struct ThPeers {
    ThMail* box[TH_MAXRANKS];      // box[r]: rank r's mailbox array (a peer mapping; box[rank] is local memory)
    int rank, world, fused, epoch; // fused = 0: the host all-reduces over NCCL instead (kept for A/B measurements);
                                   // epoch = number of the nonlinear step (tags the mailbox sequence numbers)
};

struct ThUImg { int channels; long long offset; int ptr_slot; int ndim; int dim[TH_MAXD]; long long elements; };
struct ThGroup { int ndim; int dim[TH_MAXD]; int nterms; int nnz; };
__device__ constexpr ThUImg TH_UIMG[TH_NUM_UIMG] = TH_UIMG_TABLE;
__device__ constexpr ThGroup TH_GROUPS[TH_NGROUPS] = TH_GROUP_TABLE;
__device__ constexpr long long TH_DIMS[TH_NDIMS] = TH_DIM_SIZES;
// Flat ranges [lo, hi) of a solver vector whose values the neighbours hold as ghosts: a value written at flat
// index f in segment s is also stored to dst[s][f - lo[s]] (peer memory).  vec4[s]: lo, hi and dst are aligned
// so that whole real4 chunks can be forwarded with one store.
struct ThPush { long long lo[2 * TH_NUM_UIMG]; long long hi[2 * TH_NUM_UIMG]; real* dst[2 * TH_NUM_UIMG]; int vec4[2 * TH_NUM_UIMG]; int n; int pad; };

// ---- gather schedule (graph domains / materialised Jacobians): per sparse endpoint the residual
// elements incident to every unknown element, as CSR offsets + a permutation (nullptr when the
// index array is already sorted, i.e. the identity), built by the plan from the caller's index
// arrays; per residual group the stored partial derivatives (TH_GROUP_NNZP scalars per element,
// endpoint-major) and J p per residual row.
#ifndef TH_GATHER
#define TH_GATHER 0
#endif
#ifndef TH_OWN_ENDPOINT
#define TH_OWN_ENDPOINT 1      // reads through the endpoint being walked use the walker's own element (no index load)
#endif
#if TH_GATHER
struct ThGather {
    const int* ptr[TH_NEP_S > 0 ? TH_NEP_S : 1];
    const int* perm[TH_NEP_S > 0 ? TH_NEP_S : 1];
    real* jvals[TH_NGROUPS];
    real* jp[TH_NGROUPS];
};
struct ThSpace { int nslots; int lanes; long long elements; };
struct ThSlot { int image; int channel; };
#endif

// ---- multi-GPU slab partition: the plan is compiled for the rank-local extent of the
// slowest axis INCLUDING TH_GHOST_LO / TH_GHOST_HI ghost layers that belong to the neighbouring
// ranks.  Owned elements of every unknown image form one contiguous flat range.
#ifndef TH_MULTI
#define TH_MULTI 0
#endif
#ifndef TH_GHOST_LO
#define TH_GHOST_LO 0
#define TH_GHOST_HI 0
#endif
#if TH_MULTI && defined(TH_PART_TABLE)
// graph partition (gather schedule): per dimension the owned index range [lo, size - hi) -- ghost vertices in
// front of / behind the owned vertices, foreign edges behind the owned edges -- and per unknown image the flat
// range of its owned scalars
struct ThPart { long long lo, hi; };
__device__ constexpr ThPart TH_PART[TH_NDIMS] = TH_PART_TABLE;
__device__ constexpr ThPart TH_RANGE[TH_NUM_UIMG] = TH_RANGE_TABLE;
#define TH_NRANGES TH_NUM_UIMG
__device__ constexpr long long th_range_lo(int k) { return TH_RANGE[k].lo; }
__device__ constexpr long long th_range_hi(int k) { return TH_RANGE[k].hi; }
__device__ __forceinline__ bool th_owned_slow(int) { return true; }
// replicated unknown images (TH_REP_TABLE): every rank holds and updates them in full (their part of J^T F,
// diag J^T J and J^T J p is all-reduced), and exactly one rank (TH_REP_OWNER) counts them in the dot products
#ifdef TH_REP_TABLE
#define TH_HAS_REP 1
__device__ constexpr int TH_REP[TH_NUM_UIMG] = TH_REP_TABLE;
__device__ constexpr bool th_range_counted(int k) { return !TH_REP[k] || TH_REP_OWNER; }
#endif
#elif TH_MULTI
#define TH_NRANGES TH_NUM_UIMG
__device__ constexpr long long th_range_lo(int k) {
    return TH_UIMG[k].offset + (long long)TH_GHOST_LO * (TH_UIMG[k].elements / TH_DSLOW) * TH_UIMG[k].channels;
}
__device__ constexpr long long th_range_hi(int k) {
    return TH_UIMG[k].offset + (long long)(TH_DSLOW - TH_GHOST_HI) * (TH_UIMG[k].elements / TH_DSLOW) * TH_UIMG[k].channels;
}
__device__ __forceinline__ bool th_owned_slow(int c) { return c >= TH_GHOST_LO && c < TH_DSLOW - TH_GHOST_HI; }
#else
#define TH_NRANGES 1
__device__ constexpr long long th_range_lo(int) { return 0; }
__device__ constexpr long long th_range_hi(int) { return TH_NUNK; }
__device__ __forceinline__ bool th_owned_slow(int) { return true; }
#endif

#ifndef TH_HAS_REP
#define TH_HAS_REP 0
__device__ constexpr bool th_range_counted(int) { return true; }
#endif

#if TH_TILED
// Shared-memory tile layout of the tiled operator kernel (computed by the front end):
// every staged array is a box of roww scalars x (TH_TH+2*TH_HY) x (TH_TD+2*TH_HZ); a row holds
// [padl | TH_TW elements | TH_HX elements] where padl >= TH_HX elements is the left halo padded to
// 16 bytes (TMA needs a 16-byte aligned innermost start coordinate); bases are 128-byte aligned.
// Arrays read only at the element itself are staged as a plain TH_TW x TH_TH x TH_TD box (center = 1);
// coff/croww/cbytes describe the CtC tile (LM diagonal) of an unknown image in that form.
struct ThStage { int slot; int es; int channels; int roww; int off; int padl; int center; int bytes; };
struct ThVTile { int roww; int zoff; int poff; int bytes; int padl; int coff; int croww; int cbytes; };
__device__ constexpr ThStage TH_STAGE[TH_NSTAGE > 0 ? TH_NSTAGE : 1] = TH_STAGE_TABLE;
__device__ constexpr int TH_SLOT_STAGE[TH_NPTR] = TH_SLOT_STAGE_TABLE;
__device__ constexpr ThVTile TH_VTILE[TH_NUM_UIMG] = TH_VTILE_TABLE;
#define TH_EXT_X (TH_TW + 2 * TH_HX)
#define TH_EXT_Y (TH_TH + 2 * TH_HY)
#define TH_EXT_Z (TH_TD + 2 * TH_HZ)
#define TH_TILE_THREADS (TH_TW * TH_TH * TH_TD)
// opaque 128-byte CUtensorMap (cuda.h is not available under NVRTC)
struct alignas(64) ThTensorMap { unsigned long long q[16]; };
// z: preconditioned residual; p[0], p[1]: the two search-direction buffers; p[2]: delta (for A*delta in LM);
// c: CtC (tile-only boxes); st: the staged problem images
struct ThMaps {
    ThTensorMap z[TH_NUM_UIMG];
    ThTensorMap p[3][TH_NUM_UIMG];
    ThTensorMap c[TH_NUM_UIMG];
    ThTensorMap st[TH_NSTAGE > 0 ? TH_NSTAGE : 1];
};
#endif

// ---- math (IEEE-accurate device functions; no fast-math)
__device__ __forceinline__ float th_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double th_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float th_sin(float x) { return sinf(x); }
__device__ __forceinline__ double th_sin(double x) { return sin(x); }
__device__ __forceinline__ float th_cos(float x) { return cosf(x); }
__device__ __forceinline__ double th_cos(double x) { return cos(x); }
__device__ __forceinline__ float th_tan(float x) { return tanf(x); }
__device__ __forceinline__ double th_tan(double x) { return tan(x); }
__device__ __forceinline__ float th_exp(float x) { return expf(x); }
__device__ __forceinline__ double th_exp(double x) { return exp(x); }
__device__ __forceinline__ float th_log(float x) { return logf(x); }
__device__ __forceinline__ double th_log(double x) { return log(x); }
__device__ __forceinline__ float th_abs(float x) { return x >= 0.0f ? x : -x; }     // ad.t:810
__device__ __forceinline__ double th_abs(double x) { return x >= 0.0 ? x : -x; }
__device__ __forceinline__ float th_asin(float x) { return asinf(x); }
__device__ __forceinline__ double th_asin(double x) { return asin(x); }
__device__ __forceinline__ float th_acos(float x) { return acosf(x); }
__device__ __forceinline__ double th_acos(double x) { return acos(x); }
__device__ __forceinline__ float th_atan(float x) { return atanf(x); }
__device__ __forceinline__ double th_atan(double x) { return atan(x); }
__device__ __forceinline__ float th_sinh(float x) { return sinhf(x); }
__device__ __forceinline__ double th_sinh(double x) { return sinh(x); }
__device__ __forceinline__ float th_cosh(float x) { return coshf(x); }
__device__ __forceinline__ double th_cosh(double x) { return cosh(x); }
__device__ __forceinline__ float th_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double th_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float th_pow(float x, float y) { return powf(x, y); }
__device__ __forceinline__ double th_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float th_fmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double th_fmin(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float th_fmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double th_fmax(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float th_floor(float x) { return floorf(x); }
__device__ __forceinline__ double th_floor(double x) { return floor(x); }
__device__ __forceinline__ float th_ceil(float x) { return ceilf(x); }
__device__ __forceinline__ double th_ceil(double x) { return ceil(x); }
__device__ __forceinline__ bool th_finite(float x) { return isfinite(x); }
__device__ __forceinline__ bool th_finite(double x) { return isfinite(x); }

// integer power by repeated multiplication (ad.t:737-760 genpow)
template <int N> __device__ __forceinline__ real th_powi(real a) {
    real r = (real)1;
#pragma unroll
    for (int i = 0; i < N; ++i) r = r * a;
    return r;
}

// ---- element loads: AoS pixels; 2- and 4-channel pixels are loaded as vectors
// (thallo.t:758,788-797), everything else scalar by scalar.
template <class CT, int C, int CH> struct ThLoad {
    static __device__ __forceinline__ real ld(const void* base, long long e) {
        return (real)__ldg(((const CT*)base) + e * C + CH);
    }
};
template <int CH> struct ThLoad<float, 2, CH> {
    static __device__ __forceinline__ real ld(const void* base, long long e) {
        const float2 v = __ldg(((const float2*)base) + e);
        return (real)(CH == 0 ? v.x : v.y);
    }
};
template <int CH> struct ThLoad<float, 4, CH> {
    static __device__ __forceinline__ real ld(const void* base, long long e) {
        const float4 v = __ldg(((const float4*)base) + e);
        return (real)(CH == 0 ? v.x : (CH == 1 ? v.y : (CH == 2 ? v.z : v.w)));
    }
};
template <int CH> struct ThLoad<double, 2, CH> {
    static __device__ __forceinline__ real ld(const void* base, long long e) {
        const double2 v = __ldg(((const double2*)base) + e);
        return (real)(CH == 0 ? v.x : v.y);
    }
};
