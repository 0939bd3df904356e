// thallo_b200 solver skeleton, part 2: index helpers, accessors, the scatter sink, deterministic
// reductions and the scalar helpers shared by the kernels.  Included after the generated
// per-energy device functions (namespace th) and before the generated gather functions and
// thallo_kernels.cuh.
#pragma once
#include "thallo_warp.cuh"

#define TH_BLOCK 256
// residualwise scatters through a sparse index array: reduce the contributions of the lanes of a
// warp that target the same unknown element before the atomic (thallo.t:3361-3402)
#ifndef TH_WARP_AGG
#define TH_WARP_AGG 1
#endif

// ------------------------------------------------------------------ index helpers
template <class Dom> struct ThIdx {
    int c[TH_MAXD];
    long long lin;
    __device__ __forceinline__ bool from_linear(long long l) {
        lin = l;
        const long long n = Dom::D0 * Dom::D1 * Dom::D2;
        if (l >= n) return false;
        c[0] = (int)(l % Dom::D0);
        c[1] = (int)((l / Dom::D0) % Dom::D1);
        c[2] = (int)(l / (Dom::D0 * Dom::D1));
        return true;
    }
    __device__ __forceinline__ bool from_coords(int x, int y, int z) {
        c[0] = x; c[1] = y; c[2] = z;
        lin = x + Dom::D0 * (y + Dom::D1 * (long long)z);
        return x < Dom::D0 && y < Dom::D1 && z < Dom::D2;
    }
};

// Unknownwise launch geometry: 1-D 256, 2-D 32x8, 3-D 8x8x4 threads per block
// (the reference uses 256 / 16x16 / 8x8x4, util.t:715-725; 32-wide rows coalesce better).
template <class Dom> __device__ __forceinline__ bool th_uw_index(ThIdx<Dom>& i) {
    if (Dom::ND == 1) return i.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x);
    return i.from_coords(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y * blockDim.y + threadIdx.y,
                         blockIdx.z * blockDim.z + threadIdx.z);
}

// ------------------------------------------------------------------ global-memory accessor
// Bounds-checked loads return 0 out of bounds (thallo.t:876-882); `vec` reads the
// unknown-shaped vector argument (P / Delta).
// OWNSP >= 0 (gather schedule): the sparse index array in ptr slot OWNSP is known to hold `own` at
// this element -- the unknown element whose adjacency list is being walked -- so reads through it
// need no index load and are loop-invariant across the walk (the compiler hoists them).
template <class Dom, int OWNSP = -1> struct GAcc {
    ThIdx<Dom> i;
    const real* __restrict__ v;
    long long own;
    __device__ __forceinline__ GAcc(const ThIdx<Dom>& idx, const real* vec) : i(idx), v(vec), own(0) {}
    __device__ __forceinline__ GAcc(const ThIdx<Dom>& idx, const real* vec, long long o) : i(idx), v(vec), own(o) {}

    template <int D> __device__ __forceinline__ int coord() const { return i.c[D]; }

    template <int L0, int H0, int L1, int H1, int L2, int H2> __device__ __forceinline__ bool inb() const {
        bool ok = true;
        if (L0 < 0) ok = ok && (i.c[0] + L0 >= 0);
        if (H0 > 0) ok = ok && (i.c[0] + H0 < Dom::D0);
        if (Dom::ND > 1) {
            if (L1 < 0) ok = ok && (i.c[1] + L1 >= 0);
            if (H1 > 0) ok = ok && (i.c[1] + H1 < Dom::D1);
        }
        if (Dom::ND > 2) {
            if (L2 < 0) ok = ok && (i.c[2] + L2 >= 0);
            if (H2 > 0) ok = ok && (i.c[2] + H2 < Dom::D2);
        }
        return ok;
    }
    template <int O0, int O1, int O2> __device__ __forceinline__ long long elem() const {
        return i.lin + O0 + Dom::D0 * (O1 + Dom::D1 * (long long)O2);
    }
    template <int SLOT, class CT, int C, int CH, int O0, int O1, int O2>
    __device__ __forceinline__ real img(const Params& P) const {
        if ((O0 | O1 | O2) != 0) { if (!inb<O0, O0, O1, O1, O2, O2>()) return (real)0; }
        return ThLoad<CT, C, CH>::ld(P.ptr[SLOT], elem<O0, O1, O2>());
    }
    template <int K, int CH, int O0, int O1, int O2> __device__ __forceinline__ real vec() const {
        if ((O0 | O1 | O2) != 0) { if (!inb<O0, O0, O1, O1, O2, O2>()) return (real)0; }
        return ThLoad<real, TH_UIMG[K].channels, CH>::ld(v + TH_UIMG[K].offset, elem<O0, O1, O2>());
    }
    template <int K, int CH, int O0, int O1, int O2> __device__ __forceinline__ long long ucol() const {
        if ((O0 | O1 | O2) != 0) { if (!inb<O0, O0, O1, O1, O2, O2>()) return -1; }
        return TH_UIMG[K].offset + elem<O0, O1, O2>() * TH_UIMG[K].channels + CH;
    }
    // sparse (graph) accesses: the index array lives in ptr slot SP and is indexed by this element
    template <int SP> __device__ __forceinline__ long long sidx(const Params& P) const {
        if (SP == OWNSP) return own;
        return (long long)__ldg(((const int*)P.ptr[SP]) + i.lin);
    }
    template <int SLOT, class CT, int C, int CH, int SP> __device__ __forceinline__ real simg(const Params& P) const {
        return ThLoad<CT, C, CH>::ld(P.ptr[SLOT], sidx<SP>(P));
    }
    template <int K, int CH, int SP> __device__ __forceinline__ real svec(const Params& P) const {
        return ThLoad<real, TH_UIMG[K].channels, CH>::ld(v + TH_UIMG[K].offset, sidx<SP>(P));
    }
    template <int K, int CH, int SP> __device__ __forceinline__ long long sucol(const Params& P) const {
        return TH_UIMG[K].offset + sidx<SP>(P) * TH_UIMG[K].channels + CH;
    }
    // bilinear sample, floor/ceil lerp with zero outside (thallo.t:899-907)
    template <int SLOT> __device__ __forceinline__ real samp(const Params& P, real x, real y) const {
        const real* im = (const real*)P.ptr[SLOT];
        const int x0 = (int)th_floor(x), x1 = (int)th_ceil(x);
        const int y0 = (int)th_floor(y), y1 = (int)th_ceil(y);
        const real xn = x - (real)x0, yn = y - (real)y0;
        auto get = [&](int xx, int yy) -> real {
            return (xx >= 0 && xx < Dom::D0 && yy >= 0 && yy < Dom::D1) ? __ldg(im + xx + Dom::D0 * (long long)yy) : (real)0;
        };
        const real u = ((real)1 - xn) * get(x0, y0) + xn * get(x1, y0);
        const real b = ((real)1 - xn) * get(x0, y1) + xn * get(x1, y1);
        return ((real)1 - yn) * u + yn * b;
    }
};

// ------------------------------------------------------------------ scatter sink (atomics)
// WHICH selects the target vector (0: r / Ap / Adelta, 1: preconditioner diagonal);
// out-of-bounds targets are dropped (thallo.t:3355-3390).
// Sparse targets: `prepare<SP>` (emitted once per index array by the front end, before the
// first scatter through it) loads this element's target and finds the lanes of the warp that hit
// the same target -- one MATCH per endpoint, shared by all of its channels; `sadd` then sums the
// peers' contributions with shuffles and lets the lowest peer issue the single atomic.
template <class Dom> struct GScatter {
    ThIdx<Dom> i;
    real* t0; real* t1;
    unsigned active;
    long long se[TH_NPTR];
    unsigned peers[TH_NPTR];
    __device__ __forceinline__ GScatter(const ThIdx<Dom>& idx, real* a, real* b) : i(idx), t0(a), t1(b), active(__activemask()) {}
    template <int SP> __device__ __forceinline__ void prepare(const Params& P) {
        se[SP] = (long long)__ldg(((const int*)P.ptr[SP]) + i.lin);
#if TH_WARP_AGG
        peers[SP] = th_get_peers(active, (int)se[SP]);
#endif
    }
    template <int WHICH, int K, int CH, int O0, int O1, int O2> __device__ __forceinline__ void add(real val) {
        const int x = i.c[0] + O0, y = i.c[1] + O1, z = i.c[2] + O2;
        bool ok = x >= 0 && x < Dom::D0;
        if (Dom::ND > 1) ok = ok && y >= 0 && y < Dom::D1;
        if (Dom::ND > 2) ok = ok && z >= 0 && z < Dom::D2;
        if (!ok) return;
        const long long e = i.lin + O0 + Dom::D0 * (O1 + Dom::D1 * (long long)O2);
        atomicAdd((WHICH ? t1 : t0) + TH_UIMG[K].offset + e * TH_UIMG[K].channels + CH, val);
    }
    template <int WHICH, int K, int CH, int SP> __device__ __forceinline__ void sadd(const Params& P, real val) {
        real* dst = (WHICH ? t1 : t0) + TH_UIMG[K].offset + se[SP] * TH_UIMG[K].channels + CH;
#if TH_WARP_AGG
        th_reduce_peers_atomic(active, dst, val, peers[SP]);
#else
        atomicAdd(dst, val);
#endif
    }
};

// How boundary values reach the neighbours' ghost copies (multi-GPU):
//   TH_PUSH_MODE 1 (default)  the LAST CTA of the kernel's grid reduction copies the boundary segments from local memory
//                             into the peers' memory, fences once at system scope and sends the mailbox pairs: every
//                             other CTA is exactly the single-GPU kernel (device-scope fences, no remote traffic)
//   TH_PUSH_MODE 0            every thread forwards the boundary values it computes, and EVERY CTA fences at system
//                             scope before its reduction ticket (measured: +12 us per kernel at 2048^2; and mixing
//                             device-scope and system-scope ticket fences between the CTAs of one grid hung both GPUs
//                             on B200, profiles/r02d_*)
#ifndef TH_PUSH_MODE
#define TH_PUSH_MODE 1
#endif
#define TH_FENCE_ALL (TH_MULTI && TH_PUSH_MODE == 0)
// ------------------------------------------------------------------ deterministic block/grid reduction
__device__ __forceinline__ double th_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Every thread of every block calls this with K per-thread values.  partials holds
// K * gridsize doubles.  Returns true (in all threads of exactly one block, the last to
// arrive) with tot[k] = sum over blocks in block order.
// `pushed`: this thread stored into peer memory (multi-GPU): the CTA then fences at system scope before its ticket.
template <int K> __device__ __forceinline__ bool th_grid_reduce(double (&val)[K], double (&tot)[K], double* partials,
                                                               unsigned int* ticket, bool pushed = false) {
    __shared__ double sm[K][32];
    __shared__ bool last;
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
    const unsigned int nblocks = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double w = th_warp_sum(val[k]);
        if (lane == 0) sm[k][warp] = w;
    }
    (void)pushed;
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double w = lane < nwarps ? sm[k][lane] : 0.0;
            w = th_warp_sum(w);
            if (lane == 0) partials[(size_t)k * nblocks + bid] = w;
        }
    }
    if (tid == 0) {
        // (TH_PUSH_MODE 0: this CTA's stores into peer memory must be visible to whoever sees the mailbox pair the last CTA sends)
        if (TH_FENCE_ALL) __threadfence_system(); else __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (unsigned int b = tid; b < nblocks; b += nthreads) s += __ldcg(partials + (size_t)k * nblocks + b);
        s = th_warp_sum(s);
        __syncthreads();
        if (lane == 0) sm[k][warp] = s;
        __syncthreads();
        double w = lane < nwarps ? sm[k][lane] : 0.0;
        w = th_warp_sum(w);
        tot[k] = __shfl_sync(0xffffffffu, w, 0);
    }
    if (tid == 0) *ticket = 0u;
    return true;
}

// ------------------------------------------------------------------ multi-GPU: peer stores and the in-kernel all-reduce
#if TH_MULTI
#ifndef TH_MAIL_PAIR
#define TH_MAIL_PAIR 1        // 0: values, system fence, then a separate release store of the sequence number (debug / comparison)
#endif
__device__ __forceinline__ unsigned long long th_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// one 16-byte store / load of a (value, sequence number) pair: a single transaction on NVLink, so a reader that sees
// the sequence number sees the value that travelled with it (the scheme of NCCL's LL protocols)
__device__ __forceinline__ void th_st_pair(unsigned long long* p, unsigned long long v, unsigned long long seq) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v), "l"(seq) : "memory");
}
__device__ __forceinline__ void th_ld_pair(const unsigned long long* p, unsigned long long& v, unsigned long long& seq) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(seq) : "l"(p) : "memory");
}
__device__ __forceinline__ ThMail* th_mail(ThMail* base, int kind, int parity, int src) {
    return base + (kind * 2 + parity) * TH_MAXRANKS + src;
}
// Sum K (<= 2) doubles over all ranks.  Called by every thread of warp 0 of ONE CTA per rank (the last CTA of the
// grid reduction) with this rank's partial sums in v[]; returns the totals in v[] (all lanes), added in rank order so
// that every rank holds the same bits.  Lane r serves peer r: it stores (value, seq) pairs into the peer's mailbox
// -- after a system-scope fence, so that everything this rank pushed into peer memory before (its CTAs fence their
// pushes at system scope before they take their reduction ticket) is visible to whoever sees the pair -- and then
// polls this rank's own mailbox entry of rank r.  `seq` is unique per use of the (kind, parity) slot and identical
// on all ranks.  A peer that never answers (crashed process) traps after 20 s instead of hanging the GPU.
template <int K> __device__ __forceinline__ void th_mail_allreduce(const ThPeers& R, int kind, unsigned long long seq, double (&v)[K]) {
    const int lane = (int)((threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31u);
    const int par = (int)(seq & 1ull);
    double got[2] = {0.0, 0.0};
    if (lane < R.world) {
        ThMail* out = th_mail(R.box[lane], kind, par, R.rank);
        __threadfence_system();
#if TH_MAIL_PAIR
#pragma unroll
        for (int k = 0; k < K; ++k) th_st_pair(&out->q[2 * k], (unsigned long long)__double_as_longlong(v[k]), seq);
#else
#pragma unroll
        for (int k = 0; k < K; ++k) *(volatile unsigned long long*)&out->q[2 * k] = (unsigned long long)__double_as_longlong(v[k]);
        __threadfence_system();
#pragma unroll
        for (int k = 0; k < K; ++k) *(volatile unsigned long long*)&out->q[2 * k + 1] = seq;
#endif
        const ThMail* in = th_mail(R.box[R.rank], kind, par, lane);
        const unsigned long long t0 = th_globaltimer();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            unsigned long long bits, sq;
            for (;;) {
#if TH_MAIL_PAIR
                th_ld_pair(&in->q[2 * k], bits, sq);
#else
                sq = *(volatile const unsigned long long*)&in->q[2 * k + 1];
                if (sq == seq) { __threadfence_system(); bits = *(volatile const unsigned long long*)&in->q[2 * k]; }
#endif
                if (sq == seq) break;
                if (th_globaltimer() - t0 > 20000000000ull) {
                    printf("thallo_b200: rank %d waited 20 s for rank %d (mailbox kind %d, seq %llu): peer lost\n", R.rank, lane, kind, seq);
                    __trap();
                }
            }
            got[k] = __longlong_as_double((long long)bits);
        }
        __threadfence_system();       // acquire: the peers' pushes that preceded their pairs are visible from here on
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double tot = 0.0;
        for (int r = 0; r < R.world; ++r) tot += __shfl_sync(0xffffffffu, got[k], r);
        v[k] = tot;
    }
}
// TH_PUSH_MODE 1: called by every thread of the last CTA after the grid reduction (all CTAs' values are in local memory
// and visible): boundary segments of `src` -> the neighbours' ghost copies; ends with a barrier, after which one
// system-scope fence by the thread that sends the mailbox pairs orders all of it.
__device__ __forceinline__ void th_push_segments(const ThPush& H, const real* __restrict__ src) {
#if TH_PUSH_MODE == 1
    const int tid = (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z));
    const int nthreads = (int)(blockDim.x * blockDim.y * blockDim.z);
    for (int s = 0; s < 2 * TH_NUM_UIMG; ++s) {
        if (s >= H.n) break;
        const long long lo = H.lo[s], n = H.hi[s] - H.lo[s];
        if (H.vec4[s]) {
            const real4* __restrict__ a = (const real4*)(src + lo);
            real4* __restrict__ b = (real4*)H.dst[s];
            for (long long i = tid; i < n / 4; i += nthreads) b[i] = __ldcg(a + i);
        } else {
            for (long long i = tid; i < n; i += nthreads) H.dst[s][i] = __ldcg(src + lo + i);
        }
    }
    __syncthreads();
#endif
}
// TH_PUSH_MODE 0: forward a scalar / a real4 chunk just written at flat index f / chunk i of the pushed vector
__device__ __forceinline__ bool th_push_scalar(const ThPush& H, long long f, real val) {
    bool any = false;
#if TH_PUSH_MODE == 0
#pragma unroll
    for (int s = 0; s < 2 * TH_NUM_UIMG; ++s)
        if (s < H.n && f >= H.lo[s] && f < H.hi[s]) { H.dst[s][f - H.lo[s]] = val; any = true; }
#endif
    return any;
}
__device__ __forceinline__ bool th_push_vec4(const ThPush& H, long long i, const real4& val) {
    bool any = false;
#if TH_PUSH_MODE == 0
#pragma unroll
    for (int s = 0; s < 2 * TH_NUM_UIMG; ++s) {
        if (s >= H.n || 4 * i + 3 < H.lo[s] || 4 * i >= H.hi[s]) continue;
        if (H.vec4[s]) { ((real4*)H.dst[s])[i - H.lo[s] / 4] = val; any = true; }
        else {
            any = th_push_scalar(H, 4 * i, val.x) | th_push_scalar(H, 4 * i + 1, val.y) |
                  th_push_scalar(H, 4 * i + 2, val.z) | th_push_scalar(H, 4 * i + 3, val.w);
            return any;
        }
    }
#endif
    return any;
}
#endif

__device__ __forceinline__ real th_guarded_invert(real d) {     // GuardedInvertType.CERES, gauss_newton.t:641-648
    const real s = (real)1 + th_sqrt(d);
    return (real)1 / (s * s);
}

__device__ __forceinline__ real th_alpha(const ThScalars* S) {   // safeDivideIfNotLM, gauss_newton.t:226-234
    const real num = (real)S->rz[S->it & 1], den = (real)S->aD;
#if TH_LM
    return num / den;
#else
    return den != (real)0 ? num / den : (real)0;
#endif
}
__device__ __forceinline__ real th_beta(const ThScalars* S) {
    const real num = (real)S->rz[(S->it + 1) & 1], den = (real)S->rz[S->it & 1];
#if TH_LM
    return num / den;
#else
    return den != (real)0 ? num / den : (real)0;
#endif
}

// beta as seen by th_pcg_a of iteration `it` (> 0): the previous iteration has been closed, so the
// newest numerator sits in rz[it&1] and the one before in rz[(it+1)&1].
__device__ __forceinline__ real th_beta_prev(const ThScalars* S) {
    const real num = (real)S->rz[S->it & 1], den = (real)S->rz[(S->it + 1) & 1];
#if TH_LM
    return num / den;
#else
    return den != (real)0 ? num / den : (real)0;
#endif
}

// Shared tail of both PCGInit forms: given the gradient entry g (=J^T F) and the true
// diagonal d (=diag J^T J) of one unknown scalar, produce r, preconditioner, p (and in LM
// CtC, b, SSq) and return r*p.
__device__ __forceinline__ real th_init_scalar(const Params& P, const Vecs& V, const ThPush& H, bool& pushed, long long off, real g, real d,
                                               real pre_if_off, int first_nonlinear) {
    const real r = -g;
    real pre = TH_USEPRE ? th_guarded_invert(d) : pre_if_off;
#if TH_LM
    real ssq = pre;
    if (first_nonlinear) V.SSq[off] = pre; else ssq = V.SSq[off];
    const real radius = P.trust_region_radius;
    const real ctc_raw = d / radius;
    const real mult = ((real)1 / ssq) / radius;
    const real ctc = th_fmin(th_fmax(ctc_raw, P.min_lm_diagonal * mult), P.max_lm_diagonal * mult);
    pre = (real)1 / (ctc + radius * ctc_raw);
    V.CtC[off] = ctc;
    V.b[off] = r;
#endif
    const real p = pre * r;
    V.delta[off] = (real)0;
    V.r[off] = r;
    V.pre[off] = pre;
#if TH_TILED
    V.z[off] = p;      // the first th_pcg_a of the linear solve takes p := z (beta = 0)
#else
    V.p[off] = p;
#endif
#if TH_MULTI
    pushed |= th_push_scalar(H, off, p);      // boundary elements: also into the neighbours' ghost copies
#endif
    return r * p;
}
__device__ __forceinline__ void th_zero_scalar(const Vecs& V, const ThPush& H, bool& pushed, long long off) {
#if TH_MULTI
    pushed |= th_push_scalar(H, off, (real)0);
#endif
    V.delta[off] = (real)0; V.r[off] = (real)0; V.pre[off] = (real)0; V.p[off] = (real)0;
    V.z[off] = (real)0; V.Ap[off] = (real)0;
#if TH_TILED
    V.p2[off] = (real)0;
#endif
#if TH_LM
    V.CtC[off] = (real)0; V.b[off] = (real)0; V.Adelta[off] = (real)0;
#endif
}
// sequence number of a mailbox use: (epoch of the nonlinear step, PCG iteration + 1); identical on all ranks
__device__ __forceinline__ unsigned long long th_seq(int epoch, int it1) {
    return ((unsigned long long)(unsigned int)epoch << 32) | (unsigned long long)(unsigned int)it1;
}
__device__ __forceinline__ void th_begin_linear(ThScalars* S, double rz0, int epoch) {
    S->rz[0] = rz0; S->rz[1] = 0.0; S->aD = 0.0; S->q = 0.0; S->Q0 = 0.0;
    S->it = 0; S->done = 0; S->lin_done = 0; S->epoch = epoch;
}

