// thallo_b200 solver skeleton, part 2: accessors, reductions and the GN/LM + PCG kernels.
// Included *after* the generated per-energy device functions (namespace th).
//
// Kernel-by-kernel correspondence with reference API/src/gauss_newton.t:
//   th_init_uw            PCGInit1 unknownwise :678-710 fused with the LM diagonal set-up
//                         (PCGSaveSSq :929-934, computeCtC thallo.t:3911-3937, PCGFinalizeDiagonal :936-969)
//   th_evaljtf_g<i>       PCGInit1 residualwise :998-1004
//   th_init_finish        PCGInit1_Finish :712-731 (+ the same LM diagonal set-up)
//   th_step1_uw           PCGStep1 unknownwise :734-752 / computeAdelta :755-762
//   th_applyjtj_g<i>      PCGStep1 residualwise :1006-1016 / computeAdelta :1058-1065
//   th_step1_finish       PCGStep1_Finish :774-799
//   th_pcg_b              PCGStep2 :801-843 (vectorised; in the tiled schedule it also closes the iteration)
//   th_step2_first/second :845-886
//   th_step3              PCGStep3 :889-899 + scalar hand-over :1665 + LM zeta test :1666-1686 (untiled schedules)
//   th_pcg_a              tiled schedule: PCGStep3 of the previous iteration fused with PCGStep1 of this one --
//                         p = z + beta p is formed in a TMA-staged shared-memory tile (with halo) and J^T J p
//                         is gathered from that tile; th_precompute_coef hoists the PCG-invariant
//                         transcendental sub-expressions of J out of the inner loop
//   th_update             PCGLinearUpdate :901-906     th_copy_x  savePreviousUnknowns/revertUpdate/copyUnknownwise :908-927
//   th_cost_g<i>          computeCost :1067-1079       th_modelcost_g<i>  computeModelCost :1088-1095
// Reductions: warp shuffle -> shared memory -> one partial per block -> the last block to
// finish sums the partials in a fixed order (deterministic), replacing the reference's
// one float atomic per warp (util.t:39-50, cuda_util.t:430-449) and its per-iteration memsets.
#pragma once
#include "thallo_access.cuh"

// Does this rank own domain element `i`?  Single GPU: always; slab partition: its slowest coordinate lies in the
// owned layers; graph partition: every coordinate lies in the owned range of its dimension.
template <class Dom> __device__ __forceinline__ bool th_owned(const ThIdx<Dom>& i) {
#if TH_MULTI && defined(TH_PART_TABLE)
    bool ok = i.c[0] >= TH_PART[Dom::I0].lo && i.c[0] < Dom::D0 - TH_PART[Dom::I0].hi;
    if (Dom::ND > 1) ok = ok && i.c[1] >= TH_PART[Dom::I1].lo && i.c[1] < Dom::D1 - TH_PART[Dom::I1].hi;
    if (Dom::ND > 2) ok = ok && i.c[2] >= TH_PART[Dom::I2].lo && i.c[2] < Dom::D2 - TH_PART[Dom::I2].hi;
    return ok;
#else
    return th_owned_slow(i.c[Dom::ND - 1]);
#endif
}
// flat index of an unknown scalar -> owned by this rank?
__device__ __forceinline__ bool th_flat_owned(long long f) {
#if TH_MULTI
    bool ok = false;
#pragma unroll
    for (int k = 0; k < TH_NRANGES; ++k) ok = ok || (f >= th_range_lo(k) && f < th_range_hi(k));
    return ok;
#else
    return true;
#endif
}

// flat index of an unknown scalar -> does this rank count it in the dot products?
__device__ __forceinline__ bool th_flat_counted(long long f) {
#if TH_MULTI
    bool ok = false;
#pragma unroll
    for (int k = 0; k < TH_NRANGES; ++k) ok = ok || (th_range_counted(k) && f >= th_range_lo(k) && f < th_range_hi(k));
    return ok;
#else
    return true;
#endif
}

// linear thread id within the block (blocks are 1-D, 2-D or 3-D)
__device__ __forceinline__ int th_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }

// End of PCGInit in the last block: <r, p> (all ranks' parts) opens the linear solve.  Non-fused multi-GPU plans
// publish the rank's part; the host all-reduces S->rz[0] over NCCL.
__device__ __forceinline__ void th_init_publish(ThScalars* S, double (&tot)[1], const ThPeers& R, const ThPush& H, const real* pushed_vec) {
#if TH_MULTI
    if (R.fused) th_push_segments(H, pushed_vec);       // p0 (z in the tiled schedule) of the boundary elements -> the neighbours' ghost copies
    if (R.fused && th_tid() < 32) th_mail_allreduce<1>(R, TH_MAIL_INIT, th_seq(R.epoch, 1), tot);
#endif
    if (th_tid() == 0) th_begin_linear(S, tot[0], R.epoch);
}
// <p, Ap> at the end of the operator kernels, same scheme (sequence number = PCG iteration + 1)
__device__ __forceinline__ void th_ad_publish(ThScalars* S, double (&tot)[1], const ThPeers& R, bool accumulate, bool reduce_now) {
    if (th_tid() >= 32) return;
    if (accumulate) tot[0] += S->aD;
    __syncwarp();
#if TH_MULTI
    if (R.fused && reduce_now) th_mail_allreduce<1>(R, TH_MAIL_A, th_seq(S->epoch, S->it + 1), tot);
#endif
    if (th_tid() == 0) S->aD = tot[0];
}

// ================================================================== at-output (unknownwise) kernels
#if TH_AT_OUTPUT
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_init_uw(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials, int first_nonlinear,
           const __grid_constant__ ThPeers R, const __grid_constant__ ThPush H) {
    ThIdx<th::dom_uw> idx;
    double acc[1] = {0.0};
    bool pushed = false;
    const bool inside = th_uw_index(idx);
#if TH_MULTI
    if (inside && !th_owned_slow(idx.c[th::dom_uw::ND - 1])) {
        // ghost layers: r, z, the preconditioner are the neighbours' to initialise (z arrives by their pushes); delta
        // is kept current locally (delta += alpha p with the ghost copy of p, th_pcg_b), so it restarts from zero here
#pragma unroll
        for (int k = 0; k < TH_NUM_UIMG; ++k)
#pragma unroll
            for (int ch = 0; ch < TH_UIMG[k].channels; ++ch) V.delta[TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch] = (real)0;
    }
#endif
    if (inside && th_owned_slow(idx.c[th::dom_uw::ND - 1])) {
        GAcc<th::dom_uw> a(idx, nullptr);
        const bool ex = th::exclude_u0(a, P);
        real g[TH_U], d[TH_U];
        if (!ex) th::evalJTF_uw(a, P, g, d);
        real dot = (real)0;
        int j = 0;
#pragma unroll
        for (int k = 0; k < TH_NUM_UIMG; ++k) {
#pragma unroll
            for (int ch = 0; ch < TH_UIMG[k].channels; ++ch, ++j) {
                const long long off = TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch;
                if (ex) th_zero_scalar(V, H, pushed, off);
                else dot += th_init_scalar(P, V, H, pushed, off, g[j], d[j], (real)0.25, first_nonlinear);   // d:=1 -> G(1)=0.25, gauss_newton.t:693-696
            }
        }
        acc[0] = (double)dot;
    }
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[0], pushed)) th_init_publish(S, tot, R, H, TH_TILED ? V.z : V.p);
}

// which = 0: Ap = (JtJ [+CtC]) p with alphaDenominator = <p,Ap>;  which = 1: Adelta = (JtJ [+CtC]) delta
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_step1_uw(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials, int which) {
    if (S->done) return;
    const real* __restrict__ in = which ? V.delta : V.p;
    real* __restrict__ out = which ? V.Adelta : V.Ap;
    ThIdx<th::dom_uw> idx;
    double acc[1] = {0.0};
    if (th_uw_index(idx)) {
        GAcc<th::dom_uw> a(idx, in);
        if (!th::exclude_u0(a, P)) {
            real o[TH_U];
            th::applyJTJ_uw(a, P, o);
            real dot = (real)0;
            int j = 0;
#pragma unroll
            for (int k = 0; k < TH_NUM_UIMG; ++k) {
#pragma unroll
                for (int ch = 0; ch < TH_UIMG[k].channels; ++ch, ++j) {
                    const long long off = TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch;
                    const real pv = in[off];
                    real val = o[j];
#if TH_LM
                    val += V.CtC[off] * pv;
#endif
                    out[off] = val;
                    dot += pv * val;
                }
            }
            acc[0] = (double)dot;
        }
    }
    if (which) return;
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[1])) {
        if (threadIdx.x + threadIdx.y + threadIdx.z == 0) S->aD = tot[0];
    }
}
#endif  // TH_AT_OUTPUT

// ================================================================== flat vector kernels
// Excluded unknowns hold zeros in every solver vector (written by the init kernels), so
// these streaming kernels need no mask: 0 stays 0 and contributes 0 to every dot product.

// current search direction: the tiled schedule ping-pongs p between V.p and V.p2 (iteration `it`
// reads buffer it&1 and writes buffer (it+1)&1, see th_pcg_a); the other schedules keep V.p.
__device__ __forceinline__ real* th_pcur(const Vecs& V, const ThScalars* S) {
#if TH_TILED
    return ((S->it + 1) & 1) ? V.p2 : V.p;
#else
    return V.p;
#endif
}

// End of a PCG iteration, executed by one thread of the last block to finish: the numerator
// hand-over (gauss_newton.t:1665, here just the parity of `it`), the LM zeta test (:1666-1686)
// and the progress report to the host through mapped pinned memory, which lets the host stop
// issuing iterations after an LM early exit without ever synchronising (the reference blocks on
// a 4-byte cudaMemcpy every iteration, gauss_newton.t:1667).
__device__ __forceinline__ void th_close_iteration(ThScalars* S, real q_tolerance, ThHostFlags* hf) {
    const int it = S->it;
    const int epoch = S->epoch;
    S->it = it + 1;
    S->lin_done = it + 1;
#if TH_LM
    const real Q1 = (real)S->q, Q0 = (real)S->Q0;
    if (!th_finite(Q1)) S->done = 1;
    else {
        const real zeta = (real)(it + 1) * (Q1 - Q0) / Q1;
        if (!th_finite(zeta) || zeta < q_tolerance) S->done = 1;
        else S->Q0 = (double)Q1;
    }
#endif
    if (hf) {
        // the exit decision is published before the progress counter, so a host that has seen
        // progress >= n also sees every exit taken at an iteration <= n (all ranks of a multi-GPU
        // solve must stop issuing iterations at the same point)
        if (S->done) {
            *(volatile long long*)&hf->exit_word = ((long long)epoch << 32) | (long long)(it + 1);
            __threadfence_system();
        }
        *(volatile long long*)&hf->progress = ((long long)epoch << 32) | (long long)(it + 1);
        __threadfence_system();
    }
}

// Flat loop over the owned ranges of an unknown-sized vector: 128-bit body `fv(i)` over the aligned
// middle (i counts real4's) and scalar body `fs(i)` over the at most three elements either side.
template <class FV, class FS>
__device__ __forceinline__ void th_for_owned(long long lo, long long hi, FV&& fv, FS&& fs) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long v0 = (lo + 3) / 4, v1 = hi / 4 > v0 ? hi / 4 : v0;
    for (long long i = v0 + gtid; i < v1; i += stride) fv(i);
    for (long long i = lo + gtid; i < (v0 * 4 < hi ? v0 * 4 : hi); i += stride) fs(i);
    for (long long i = v1 * 4 + gtid; i < hi; i += stride) fs(i);
}
// PCGStep2: alpha = rz/aD; delta += alpha p; r -= alpha Ap; z = M r; <z,r>; LM: q = 1/2 <delta, r + b>.
// Pure streaming: 128-bit loads of every operand first, then the stores (the vectors never alias,
// but the compiler cannot know that through the pointer table).
#define TH_B_LANE(c)                                                             \
    {                                                                            \
        const real dn_ = dl.c + alpha * p.c;                                     \
        const real rn_ = r.c - alpha * ap.c;                                     \
        const real z_ = TH_USEPRE ? pre.c * rn_ : rn_;                           \
        dl.c = dn_; r.c = rn_; zz.c = z_;                                        \
        accr[0] += (double)(z_ * rn_);                                           \
        if (TH_LM) accr[1] += (double)((real)0.5 * (dn_ * (rn_ + bb.c)));        \
    }
// End of step 2 in the last block: publish the two sums.  Multi-GPU plans publish this rank's
// partial sums instead; the host all-reduces them over NCCL and th_mg_close finishes the iteration.
// Called by every thread of the last block.  Fused multi-GPU plans all-reduce the two sums right here over the
// peers' mailboxes (the z boundary layers this rank pushed are ordered before its flag, so a rank that has the
// totals also has its ghost copies of z) and close the iteration like the tiled single-GPU schedule does.
__device__ __forceinline__ void th_step2_publish(ThScalars* S, double (&tot)[2], real q_tolerance, ThHostFlags* hf, const ThPeers& R,
                                                 const ThPush& H, const real* z) {
#if TH_MULTI
    // R.fused == 2 (default) / 0: this rank's parts only; th_push_close (fused) or NCCL + th_mg_close (comparison path)
    // finish the iteration.  R.fused == 1: the last CTA pushes the boundary layers itself (fine for thin boundaries).
    if (R.fused != 1) {
        if (th_tid() == 0) { S->red[0] = tot[0]; S->red[1] = tot[1]; }
        return;
    }
    th_push_segments(H, z);
    if (th_tid() >= 32) return;
    th_mail_allreduce<2>(R, TH_MAIL_B, th_seq(S->epoch, S->it + 1), tot);
#else
    if (th_tid() >= 32) return;
#endif
    if (th_tid() == 0) {
        S->rz[(S->it + 1) & 1] = tot[0]; S->q = tot[1];
#if TH_TILED || TH_MULTI
        th_close_iteration(S, q_tolerance, hf);
#endif
    }
}
#if TH_MULTI
// Fused multi-GPU end of PCGStep2: a few CTAs copy the boundary layers of z into the neighbours' ghost layers (a single
// CTA takes ~100 us for the 1.2 MB faces of a 160^3 slab, measured: profiles/r02h_n8_*_last_cta_push.*), EVERY CTA fences
// at system scope before its ticket, and the last one all-reduces this rank's <z,r> and q with the peers' over the
// mailboxes and closes the iteration: whoever has the totals also has its ghost copies of z.
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_push_close(const __grid_constant__ Vecs V, ThScalars* S, real q_tolerance, ThHostFlags* hf,
              const __grid_constant__ ThPeers R, const __grid_constant__ ThPush H) {
    if (S->done) return;
    const real* __restrict__ src = V.z;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < 2 * TH_NUM_UIMG; ++s) {
        if (s >= H.n) break;
        const long long lo = H.lo[s], n = H.hi[s] - H.lo[s];
        if (H.vec4[s]) {
            const real4* __restrict__ a = (const real4*)(src + lo);
            real4* __restrict__ b = (real4*)H.dst[s];
            for (long long i = gtid; i < n / 4; i += stride) b[i] = __ldcg(a + i);
        } else {
            for (long long i = gtid; i < n; i += stride) H.dst[s][i] = __ldcg(src + lo + i);
        }
    }
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        last = atomicAdd(&S->ticket[5], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double tot[2] = {S->red[0], S->red[1]};
        th_mail_allreduce<2>(R, TH_MAIL_B, th_seq(S->epoch, S->it + 1), tot);
        if (threadIdx.x == 0) {
            S->ticket[5] = 0u;
            S->rz[(S->it + 1) & 1] = tot[0]; S->q = tot[1];
            th_close_iteration(S, q_tolerance, hf);
        }
    }
}
extern "C" __global__ void th_mg_close(ThScalars* S, real q_tolerance, ThHostFlags* hf) {
    if (S->done) return;
    S->rz[(S->it + 1) & 1] = S->red[0]; S->q = S->red[1];
    th_close_iteration(S, q_tolerance, hf);
}
// halo push: copy contiguous segments (boundary layers of every unknown image) into the
// neighbours' ghost layers over NVLink peer mappings
struct ThSegs { const real* src[2 * TH_NUM_UIMG]; real* dst[2 * TH_NUM_UIMG]; long long count[2 * TH_NUM_UIMG]; };
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_halo_push(const __grid_constant__ ThSegs G, int nseg, const ThScalars* S, int check_done) {
    if (check_done && S->done) return;
    for (int s = 0; s < nseg; ++s) {
        const real* __restrict__ src = G.src[s];
        real* __restrict__ dst = G.dst[s];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.count[s]; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
    }
}
#endif

extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_pcg_b(const __grid_constant__ Vecs V, ThScalars* S, double* partials, real q_tolerance, ThHostFlags* hf,
         const __grid_constant__ ThPeers R, const __grid_constant__ ThPush H) {
    if (S->done) return;
    const real alpha = th_alpha(S);
    const real* __restrict__ pp = th_pcur(V, S);
    real* __restrict__ vd = V.delta;
    real* __restrict__ vr = V.r;
    real* __restrict__ vz = V.z;
    const real* __restrict__ vap = V.Ap;
    const real* __restrict__ vpre = V.pre;
    const real* __restrict__ vb = V.b;
    double acc[2] = {0.0, 0.0}, accr[2] = {0.0, 0.0};
    bool pushed = false;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    auto scalar = [&](long long i) {
        const real pv = pp[i];
        const real dn = vd[i] + alpha * pv;
        const real rn = vr[i] - alpha * vap[i];
        const real zv = TH_USEPRE ? vpre[i] * rn : rn;
        vd[i] = dn; vr[i] = rn; vz[i] = zv;
#if TH_MULTI
        pushed |= th_push_scalar(H, i, zv);
#endif
        accr[0] += (double)(zv * rn);
        if (TH_LM) accr[1] += (double)((real)0.5 * (dn * (rn + vb[i])));
    };
#if TH_MULTI
    // ghost copies of delta: delta += alpha p with the locally maintained ghost copy of p (the same bits as the
    // owner's), so A*delta, the model cost and the update never need an exchange of delta
#pragma unroll
    for (int k = 0; k < TH_NUM_UIMG; ++k) {
        const long long ilo = TH_UIMG[k].offset, ihi = ilo + TH_UIMG[k].elements * TH_UIMG[k].channels;
        for (long long i = ilo + gtid; i < th_range_lo(k); i += stride) vd[i] = vd[i] + alpha * pp[i];
        for (long long i = th_range_hi(k) + gtid; i < ihi; i += stride) vd[i] = vd[i] + alpha * pp[i];
    }
#endif
#pragma unroll
    for (int k = 0; k < TH_NRANGES; ++k) {          // owned flat ranges (one range = everything on a single GPU)
        const long long lo = th_range_lo(k), hi = th_range_hi(k);
        const long long v0 = (lo + 3) / 4, v1 = hi / 4 > v0 ? hi / 4 : v0;
        for (long long i = v0 + gtid; i < v1; i += stride) {
            const real4 p = ((const real4*)pp)[i];
            real4 dl = ((const real4*)vd)[i];
            real4 r = ((const real4*)vr)[i];
            const real4 ap = ((const real4*)vap)[i];
            real4 pre, bb, zz;
            if (TH_USEPRE) pre = ((const real4*)vpre)[i];
            if (TH_LM) bb = ((const real4*)vb)[i];
            TH_B_LANE(x) TH_B_LANE(y) TH_B_LANE(z) TH_B_LANE(w)
            ((real4*)vd)[i] = dl;
            ((real4*)vr)[i] = r;
            ((real4*)vz)[i] = zz;
#if TH_MULTI
            pushed |= th_push_vec4(H, i, zz);
#endif
        }
        for (long long i = lo + gtid; i < (v0 * 4 < hi ? v0 * 4 : hi); i += stride) scalar(i);
        for (long long i = v1 * 4 + gtid; i < hi; i += stride) scalar(i);
        if (th_range_counted(k)) { acc[0] += accr[0]; acc[1] += accr[1]; }     // a replicated range counts on one rank only
        accr[0] = accr[1] = 0.0;
    }
    double tot[2];
    if (th_grid_reduce<2>(acc, tot, partials, &S->ticket[2], pushed)) th_step2_publish(S, tot, q_tolerance, hf, R, H, V.z);
}

extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_step2_first(const __grid_constant__ Vecs V, ThScalars* S) {
    if (S->done) return;
    const real alpha = th_alpha(S);
    const real* __restrict__ pp = th_pcur(V, S);
    real* __restrict__ vd = V.delta;
    // (ghost copies of delta included: they are maintained locally, see th_pcg_b)
    {
        th_for_owned(0, TH_NUNK,
            [&](long long i) {
                const real4 p = ((const real4*)pp)[i];
                real4 d = ((const real4*)vd)[i];
                d.x = d.x + alpha * p.x; d.y = d.y + alpha * p.y; d.z = d.z + alpha * p.z; d.w = d.w + alpha * p.w;
                ((real4*)vd)[i] = d;
            },
            [&](long long i) { vd[i] = vd[i] + alpha * pp[i]; });
    }
}

// r = b - A delta; add_ctc: A delta still lacks the CtC*delta term (residualwise / materialized schedules)
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_step2_second(const __grid_constant__ Vecs V, ThScalars* S, double* partials, int add_ctc, real q_tolerance, ThHostFlags* hf,
                const __grid_constant__ ThPeers R, const __grid_constant__ ThPush H) {
    if (S->done) return;
    double acc[2] = {0.0, 0.0};
    const real* __restrict__ vdl = V.delta;
    const real* __restrict__ vad = V.Adelta;
    const real* __restrict__ vctc = V.CtC;
    const real* __restrict__ vb = V.b;
    const real* __restrict__ vpre = V.pre;
    real* __restrict__ vr = V.r;
    real* __restrict__ vz = V.z;
    double accr[2] = {0.0, 0.0};
    bool pushed = false;
    auto lane = [&](real delta, real Ax, real ctc, real b, real pre, real& r, real& z) {
        if (add_ctc) Ax += ctc * delta;
        r = b - Ax;
        z = TH_USEPRE ? pre * r : r;
        accr[0] += (double)(z * r);
        accr[1] += (double)((real)0.5 * (delta * (r + b)));
    };
#pragma unroll
    for (int k = 0; k < TH_NRANGES; ++k) {
        th_for_owned(th_range_lo(k), th_range_hi(k),
            [&](long long i) {
                const real4 dl = ((const real4*)vdl)[i];
                const real4 ad = ((const real4*)vad)[i];
                const real4 bb = ((const real4*)vb)[i];
                real4 ct = dl, pr = dl, r, z;
                if (add_ctc) ct = ((const real4*)vctc)[i];
                if (TH_USEPRE) pr = ((const real4*)vpre)[i];
                lane(dl.x, ad.x, ct.x, bb.x, pr.x, r.x, z.x);
                lane(dl.y, ad.y, ct.y, bb.y, pr.y, r.y, z.y);
                lane(dl.z, ad.z, ct.z, bb.z, pr.z, r.z, z.z);
                lane(dl.w, ad.w, ct.w, bb.w, pr.w, r.w, z.w);
                ((real4*)vr)[i] = r;
                ((real4*)vz)[i] = z;
#if TH_MULTI
                pushed |= th_push_vec4(H, i, z);
#endif
            },
            [&](long long i) {
                real r, z;
                lane(vdl[i], vad[i], add_ctc ? vctc[i] : (real)0, vb[i], TH_USEPRE ? vpre[i] : (real)1, r, z);
                vr[i] = r;
                vz[i] = z;
#if TH_MULTI
                pushed |= th_push_scalar(H, i, z);
#endif
            });
        if (th_range_counted(k)) { acc[0] += accr[0]; acc[1] += accr[1]; }     // a replicated range counts on one rank only
        accr[0] = accr[1] = 0.0;
    }
    double tot[2];
    if (th_grid_reduce<2>(acc, tot, partials, &S->ticket[2], pushed)) th_step2_publish(S, tot, q_tolerance, hf, R, H, V.z);
}

// PCGStep3 of the untiled schedules: beta = rz_new/rz_old; p = z + beta p; closes the iteration.
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_step3(const __grid_constant__ Vecs V, ThScalars* S, real q_tolerance, ThHostFlags* hf) {
    if (S->done) return;
#if TH_MULTI
    const real beta = th_beta_prev(S);       // th_pcg_b (fused) / th_mg_close has already closed the iteration (after the all-reduce of <z,r>)
#else
    const real beta = th_beta(S);
#endif
    const real* __restrict__ vz = V.z;
    real* __restrict__ vp = V.p;
    th_for_owned(0, TH_NUNK,
        [&](long long i) {
            const real4 z = ((const real4*)vz)[i];
            real4 p = ((const real4*)vp)[i];
            p.x = z.x + beta * p.x; p.y = z.y + beta * p.y; p.z = z.z + beta * p.z; p.w = z.w + beta * p.w;
            ((real4*)vp)[i] = p;
        },
        [&](long long i) { vp[i] = vz[i] + beta * vp[i]; });
#if !TH_MULTI
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(&S->ticket[3], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        S->ticket[3] = 0u;
        th_close_iteration(S, q_tolerance, hf);
    }
#endif
}

// ================================================================== hoisted invariants
#if TH_AT_OUTPUT && TH_NCOEF > 0
// Evaluated once per nonlinear iteration (J is constant across the PCG iterations of a step).
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_precompute_coef(const __grid_constant__ Params P) {
    ThIdx<th::dom_uw> idx;
    if (th_uw_index(idx)) {
        GAcc<th::dom_uw> a(idx, nullptr);
        real c[TH_NCOEF];
        th::coef_uw(a, P, c);
        real* __restrict__ out = (real*)P.ptr[TH_COEF_SLOT];
#pragma unroll
        for (int i = 0; i < TH_NCOEF; ++i) out[idx.lin * TH_NCOEF + i] = c[i];
    }
}
#endif

// ================================================================== tiled operator kernel (2-D / 3-D image domains)
#if TH_TILED
#ifndef TH_PCG_A_MINB
#define TH_PCG_A_MINB 3
#endif
#ifndef TH_PIPE
#define TH_PIPE 2
#endif
__device__ __forceinline__ unsigned th_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void th_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(th_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void th_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(th_smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the warp in hardware for a bounded time and is simply retried.  A __nanosleep
// backoff between attempts (TH_WAIT_SLEEP_NS > 0) was measured and loses on every tiled workload: a late
// wake-up costs more than the issue slots the spin takes (th_pcg_a, 64 ns vs none: image_warping 2048^2
// 0.0731 / 0.0689 ms, volumetric 160^3 0.2360 / 0.2270 ms, shape_from_shading 4096^2 0.3891 / 0.3861 ms;
// profiles/r01j_sweep.txt, r01k_sweep.txt).
#ifndef TH_WAIT_SLEEP_NS
#define TH_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ bool th_mbar_try(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(th_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void th_mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!th_mbar_try(bar, parity)) {
        if (TH_WAIT_SLEEP_NS > 0) __nanosleep(TH_WAIT_SLEEP_NS);
    }
}
// TMA: one bulk tensor copy of a whole (tile + halo) box into shared memory; coordinates may be
// negative or run past the image, the hardware fills out-of-bounds elements with zeros -- exactly
// the reference's out-of-bounds load semantics (thallo.t:876-882).
__device__ __forceinline__ void th_tma_load(void* dst, const ThTensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
#if TH_UW_NDIM == 2
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(th_smem_u32(dst)), "l"((unsigned long long)map), "r"(th_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(th_smem_u32(dst)), "l"((unsigned long long)map), "r"(th_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
#endif
}

__device__ __forceinline__ void th_mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(th_smem_u32(bar)) : "memory");
}

// Fallback loader (rows not 16-byte aligned, so no tensor map can describe the image): the block
// fills the same box cooperatively with bounds-checked loads.  center: tile-only box, no halo.
template <class T>
__device__ __forceinline__ void th_tile_load(T* __restrict__ dst, const T* __restrict__ src, int channels, int roww, int padl,
                                             int center, int x0, int y0, int z0, int tid) {
    const int ey = center ? TH_TH : TH_EXT_Y, ez = center ? TH_TD : TH_EXT_Z;
    const int hy = center ? 0 : TH_HY, hz = center ? 0 : TH_HZ;
    const int total = roww * ey * ez;
    const long long W = th::dom_uw::D0 * channels, H = th::dom_uw::D1, D = th::dom_uw::D2;
    for (int e = tid; e < total; e += TH_TILE_THREADS) {
        const int c = e % roww, yy = (e / roww) % ey, zz = e / (roww * ey);
        const long long gx = (long long)x0 * channels - padl + c, gy = y0 - hy + yy, gz = z0 - hz + zz;
        const bool ok = gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < D;
        dst[e] = ok ? src[gx + W * (gy + H * gz)] : (T)0;
    }
}
__device__ __forceinline__ void th_tile_load_es(void* dst, const void* src, int es, int channels, int roww, int padl, int center,
                                                int x0, int y0, int z0, int tid) {
    if (es == 1) th_tile_load((unsigned char*)dst, (const unsigned char*)src, channels, roww, padl, center, x0, y0, z0, tid);
    else if (es == 4) th_tile_load((unsigned int*)dst, (const unsigned int*)src, channels, roww, padl, center, x0, y0, z0, tid);
    else th_tile_load((unsigned long long*)dst, (const unsigned long long*)src, channels, roww, padl, center, x0, y0, z0, tid);
}

// Accessor over the staged tiles: stencil taps of the vector argument and of every staged image
// come from shared memory (zero outside the image, courtesy of the loader); anything not staged
// is read at the element itself straight from global memory.  With UPD the vector argument is the
// new search direction p = z + beta p_old, formed on the fly from the z and p_old tiles
// (PCGStep3 fused into the operator; fma() so that every tap and the stored p_new agree bit for bit).
// EDGE = false: the tile lies far enough inside the domain (TH_INB_R* elements) that every bounds predicate of the
// generated code is true; they then fold away at compile time together with the selects they guard.
template <class Dom, bool UPD, bool EDGE = true> struct TAcc {
    ThIdx<Dom> i;
    const unsigned char* sm;
    int tx, ty, tz;
    real beta;
    __device__ __forceinline__ TAcc(const ThIdx<Dom>& idx, const unsigned char* s, int x, int y, int z, real b)
        : i(idx), sm(s), tx(x), ty(y), tz(z), beta(b) {}
    template <int D> __device__ __forceinline__ int coord() const { return i.c[D]; }
    template <int L0, int H0, int L1, int H1, int L2, int H2> __device__ __forceinline__ bool inb() const {
        if (!EDGE) return true;
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
    template <int O0, int O1, int O2> __device__ __forceinline__ int tile_elem(int roww, int padl, int channels) const {
        return ((tz + TH_HZ + O2) * TH_EXT_Y + (ty + TH_HY + O1)) * roww + padl + (tx + O0) * channels;
    }
    __device__ __forceinline__ int center_elem(int roww, int channels) const { return (tz * TH_TH + ty) * roww + tx * channels; }
    template <int SLOT, class CT, int C, int CH, int O0, int O1, int O2>
    __device__ __forceinline__ real img(const Params& P) const {
        constexpr int s = TH_SLOT_STAGE[SLOT];
        if constexpr (s >= 0) {
            constexpr ThStage st = TH_STAGE[s >= 0 ? s : 0];
            const CT* t = (const CT*)(sm + st.off);
            if constexpr (st.center != 0) {
                static_assert((O0 | O1 | O2) == 0, "image staged without halo is read at an offset");
                return (real)t[center_elem(st.roww, C) + CH];
            } else {
                return (real)t[tile_elem<O0, O1, O2>(st.roww, st.padl, C) + CH];
            }
        } else {
            static_assert((O0 | O1 | O2) == 0, "image read at an offset must be staged");
            return ThLoad<CT, C, CH>::ld(P.ptr[SLOT], i.lin);
        }
    }
    __device__ __forceinline__ real vec_at(int k, int e) const {
        const real pv = ((const real*)(sm + TH_VTILE[k].poff))[e];
        if (UPD) return fma(beta, pv, ((const real*)(sm + TH_VTILE[k].zoff))[e]);
        return pv;
    }
    template <int K, int CH, int O0, int O1, int O2> __device__ __forceinline__ real vec() const {
        return vec_at(K, tile_elem<O0, O1, O2>(TH_VTILE[K].roww, TH_VTILE[K].padl, TH_UIMG[K].channels) + CH);
    }
    template <int SLOT> __device__ __forceinline__ real samp(const Params& P, real x, real y) const {
        GAcc<Dom> g(i, nullptr);
        return g.template samp<SLOT>(P, x, y);
    }
};

// mode 0: p_new = z + beta p_old (beta = 0 on the first iteration), Ap = (JtJ [+CtC]) p_new,
//         alphaDenominator = <p_new, Ap>; p_old is read from buffer it&1, p_new written to (it+1)&1
//         (neighbouring tiles still read p_old in their halos).
// mode 1: Adelta = (JtJ [+CtC]) delta   (LM residual reset, gauss_newton.t:755-762)
//
// Persistent CTAs (the host launches SMs x resident-CTAs-per-SM of them) walk the tile list with a
// TH_PIPE-stage shared-memory pipeline (two stages when they fit beside TH_PCG_A_MINB resident
// CTAs, else one): while tile i is processed, the TMA unit already fills the other stage with tile
// i+1 (every array the operator reads, so no thread issues a global load), a stage is handed
// back through an mbarrier the warps arrive on, and there is a single grid reduction per CTA at
// the very end.
#define TH_NTX ((int)((th::dom_uw::D0 + TH_TW - 1) / TH_TW))
#define TH_NTY ((int)((th::dom_uw::D1 + TH_TH - 1) / TH_TH))
#define TH_NTZ ((int)((th::dom_uw::D2 + TH_TD - 1) / TH_TD))
#define TH_NTILES (TH_NTX * TH_NTY * TH_NTZ)
#define TH_CTC_STAGED (TH_LM && TH_STAGE_CTC)

__device__ __forceinline__ void th_tile_origin(int t, int& x0, int& y0, int& z0) {
    x0 = (t % TH_NTX) * TH_TW;
    y0 = ((t / TH_NTX) % TH_NTY) * TH_TH;
    z0 = (t / (TH_NTX * TH_NTY)) * TH_TD;
}

// all bulk copies of one tile, issued by one thread
__device__ __forceinline__ void th_tile_issue(unsigned char* sm, unsigned long long* bar, const ThMaps& M, int t, int mode, int it) {
    int x0, y0, z0;
    th_tile_origin(t, x0, y0, z0);
    const bool upd = mode == 0 && it > 0;
    const int psrc = mode ? 2 : (it & 1);
    unsigned bytes = 0;
#pragma unroll
    for (int k = 0; k < TH_NUM_UIMG; ++k) bytes += (unsigned)TH_VTILE[k].bytes * (upd ? 2u : 1u) + (TH_CTC_STAGED ? (unsigned)TH_VTILE[k].cbytes : 0u);
#pragma unroll
    for (int s = 0; s < TH_NSTAGE; ++s) bytes += (unsigned)TH_STAGE[s].bytes;
    th_mbar_expect_tx(bar, bytes);
#pragma unroll
    for (int k = 0; k < TH_NUM_UIMG; ++k) {
        const int c0 = x0 * TH_UIMG[k].channels - TH_VTILE[k].padl;
        if (mode == 0 && it == 0) th_tma_load(sm + TH_VTILE[k].poff, &M.z[k], bar, c0, y0 - TH_HY, z0 - TH_HZ);
        else {
            if (upd) th_tma_load(sm + TH_VTILE[k].zoff, &M.z[k], bar, c0, y0 - TH_HY, z0 - TH_HZ);
            th_tma_load(sm + TH_VTILE[k].poff, &M.p[psrc][k], bar, c0, y0 - TH_HY, z0 - TH_HZ);
        }
        if (TH_CTC_STAGED) th_tma_load(sm + TH_VTILE[k].coff, &M.c[k], bar, x0 * TH_UIMG[k].channels, y0, z0);
    }
#pragma unroll
    for (int s = 0; s < TH_NSTAGE; ++s) {
        if (TH_STAGE[s].center) th_tma_load(sm + TH_STAGE[s].off, &M.st[s], bar, x0 * TH_STAGE[s].channels, y0, z0);
        else th_tma_load(sm + TH_STAGE[s].off, &M.st[s], bar, x0 * TH_STAGE[s].channels - TH_STAGE[s].padl, y0 - TH_HY, z0 - TH_HZ);
    }
}

// the same fill with cooperative bounds-checked loads (synchronous; all threads)
__device__ __forceinline__ void th_tile_fill(unsigned char* sm, const Params& P, const Vecs& V, int t, int mode, int it, int tid) {
    int x0, y0, z0;
    th_tile_origin(t, x0, y0, z0);
    const bool upd = mode == 0 && it > 0;
    const real* psrcv = mode ? V.delta : ((it & 1) ? V.p2 : V.p);
#pragma unroll
    for (int k = 0; k < TH_NUM_UIMG; ++k) {
        const int ch = TH_UIMG[k].channels;
        if (mode == 0 && it == 0) th_tile_load((real*)(sm + TH_VTILE[k].poff), V.z + TH_UIMG[k].offset, ch, TH_VTILE[k].roww, TH_VTILE[k].padl, 0, x0, y0, z0, tid);
        else {
            if (upd) th_tile_load((real*)(sm + TH_VTILE[k].zoff), V.z + TH_UIMG[k].offset, ch, TH_VTILE[k].roww, TH_VTILE[k].padl, 0, x0, y0, z0, tid);
            th_tile_load((real*)(sm + TH_VTILE[k].poff), psrcv + TH_UIMG[k].offset, ch, TH_VTILE[k].roww, TH_VTILE[k].padl, 0, x0, y0, z0, tid);
        }
        if (TH_CTC_STAGED) th_tile_load((real*)(sm + TH_VTILE[k].coff), V.CtC + TH_UIMG[k].offset, ch, TH_VTILE[k].croww, 0, 1, x0, y0, z0, tid);
    }
#pragma unroll
    for (int s = 0; s < TH_NSTAGE; ++s)
        th_tile_load_es(sm + TH_STAGE[s].off, P.ptr[TH_STAGE[s].slot], TH_STAGE[s].es, TH_STAGE[s].channels, TH_STAGE[s].roww,
                        TH_STAGE[s].padl, TH_STAGE[s].center, x0, y0, z0, tid);
}

// operator applied to one tile from shared-memory stage `sm`; returns this thread's <p, Ap> contribution
#ifndef TH_INB_RX
#define TH_INB_RX 1000000
#define TH_INB_RY 1000000
#define TH_INB_RZ 1000000
#endif
// does any bounds predicate of the generated code possibly fail on this tile (or on the positions around it)?
__device__ __forceinline__ bool th_tile_is_edge(int t) {
    int x0, y0, z0;
    th_tile_origin(t, x0, y0, z0);
    bool e = x0 - TH_INB_RX < 0 || x0 + TH_TW + TH_INB_RX > th::dom_uw::D0;
    e = e || y0 - TH_INB_RY < 0 || y0 + TH_TH + TH_INB_RY > th::dom_uw::D1;
    if (th::dom_uw::ND > 2) e = e || z0 - TH_INB_RZ < 0 || z0 + TH_TD + TH_INB_RZ > th::dom_uw::D2;
    return e;
}
template <bool UPD, bool EDGE>
__device__ __forceinline__ real th_tile_apply(const unsigned char* sm, const Params& P, const Vecs& V, int t, int mode, real beta,
                                              real* __restrict__ out, real* __restrict__ pnew, int tx, int ty, int tz) {
    int x0, y0, z0;
    th_tile_origin(t, x0, y0, z0);
    ThIdx<th::dom_uw> idx;
    real dot = (real)0;
    if (idx.from_coords(x0 + tx, y0 + ty, z0 + tz)) {
        TAcc<th::dom_uw, UPD, EDGE> a(idx, sm, tx, ty, tz, beta);
#if TH_MULTI
        if (!th_owned_slow(idx.c[th::dom_uw::ND - 1])) {
            // ghost layer: the owner computes Ap there; this rank only keeps its copy of the search
            // direction current (same z, same p_old, same beta -> the same bits as the owner's)
            if (mode == 0) {
#pragma unroll
                for (int k = 0; k < TH_NUM_UIMG; ++k) {
                    const int te = a.template tile_elem<0, 0, 0>(TH_VTILE[k].roww, TH_VTILE[k].padl, TH_UIMG[k].channels);
#pragma unroll
                    for (int ch = 0; ch < TH_UIMG[k].channels; ++ch)
                        pnew[TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch] = a.vec_at(k, te + ch);
                }
            }
            return dot;
        }
#endif
        if (!th::exclude_u0(a, P)) {
            real o[TH_U];
            th::applyJTJ_uw(a, P, o);
            int j = 0;
#pragma unroll
            for (int k = 0; k < TH_NUM_UIMG; ++k) {
                const int te = a.template tile_elem<0, 0, 0>(TH_VTILE[k].roww, TH_VTILE[k].padl, TH_UIMG[k].channels);
#pragma unroll
                for (int ch = 0; ch < TH_UIMG[k].channels; ++ch, ++j) {
                    const long long off = TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch;
                    const real pv = a.vec_at(k, te + ch);
                    real val = o[j];
#if TH_LM
#if TH_STAGE_CTC
                    val += ((const real*)(sm + TH_VTILE[k].coff))[a.center_elem(TH_VTILE[k].croww, TH_UIMG[k].channels) + ch] * pv;
#else
                    val += V.CtC[off] * pv;
#endif
#endif
                    out[off] = val;
                    if (mode == 0) pnew[off] = pv;
                    dot += pv * val;
                }
            }
        }
    }
    return dot;
}

#ifndef TH_TWO_PHASE
#define TH_TWO_PHASE 0
#endif
#if TH_TWO_PHASE
// Two-phase form of the tile operator (front end: Generator.gen_two_phase): J p of every residual term is formed ONCE
// per residual position into shared-memory planes (phase 1: the tile's own positions, and the positions around the
// tile whose residuals reach into it -- only the term classes that do), then every unknown multiplies its partial
// derivative of each residual instance with the stored value (phase 2).  The Jt[Jp] schedule of the reference
// (PCGStep1_J / PCGStep1_Jt, gauss_newton.t:1027-1047) with J p living in shared memory instead of HBM.
#define TH_JP_BX (TH_TW + 2 * TH_JP_PHX)
#define TH_JP_BY (TH_TH + 2 * TH_JP_PHY)
#define TH_JP_BZ (TH_TD + 2 * TH_JP_PHZ)
#define TH_JP_NBOX (TH_JP_BX * TH_JP_BY * TH_JP_BZ)
#define TH_JP_BYTES (TH_JP_NT * TH_JP_NBOX * (int)sizeof(real))
// halo positions relative to the tile origin: (dx + 8) | (dy + 8) << 8 | (dz + 8) << 16 | class mask << 24
__device__ const unsigned int TH_JP_POS[TH_JP_NHALO > 0 ? TH_JP_NHALO : 1] = TH_JP_POS_TABLE;
struct ThJp {
    real* buf;
    int tx, ty, tz;
    static __device__ __forceinline__ int box(int x, int y, int z) {
        return ((z + TH_JP_PHZ) * TH_JP_BY + (y + TH_JP_PHY)) * TH_JP_BX + (x + TH_JP_PHX);
    }
    template <int T, int S0, int S1, int S2> __device__ __forceinline__ real at() const {
        return buf[T * TH_JP_NBOX + box(tx + S0, ty + S1, tz + S2)];
    }
    template <int T> __device__ __forceinline__ void put(real v) { buf[T * TH_JP_NBOX + box(tx, ty, tz)] = v; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};
template <bool UPD, bool EDGE>
__device__ __forceinline__ real th_tile_apply2(const unsigned char* sm, real* jpbuf, const Params& P, const Vecs& V, int t, int mode, real beta,
                                               real* __restrict__ out, real* __restrict__ pnew, int tx, int ty, int tz) {
    int x0, y0, z0;
    th_tile_origin(t, x0, y0, z0);
    const int tid = tx + TH_TW * (ty + TH_TH * tz);
    // one set of planes: wait until the previous tile's phase 2 has read them.  Two sets alternate per tile: a warp
    // that starts phase 1 of tile k+1 has passed the barrier of tile k, which every warp reaches only after its own
    // phase 2 of tile k-1 -- the set being overwritten is no longer read.
    if (TH_JP_BUFS == 1) __syncthreads();
    // phase 1, positions around the tile
    for (int q = tid; q < TH_JP_NHALO; q += TH_TILE_THREADS) {
        const unsigned int e = __ldg(&TH_JP_POS[q]);
        const int dx = (int)(e & 255u) - 8, dy = (int)((e >> 8) & 255u) - 8, dz = (int)((e >> 16) & 255u) - 8;
        real jp[TH_JP_NT];
#pragma unroll
        for (int k = 0; k < TH_JP_NT; ++k) jp[k] = (real)0;
        ThIdx<th::dom_uw> idx;
        if (x0 + dx >= 0 && y0 + dy >= 0 && z0 + dz >= 0 && idx.from_coords(x0 + dx, y0 + dy, z0 + dz)) {
            TAcc<th::dom_uw, UPD, EDGE> a(idx, sm, dx, dy, dz, beta);
            th::applyJ_halo(a, P, e >> 24, jp);
        }
        const int b = ThJp::box(dx, dy, dz);
#pragma unroll
        for (int k = 0; k < TH_JP_NT; ++k) jpbuf[k * TH_JP_NBOX + b] = jp[k];
    }
    // phase 1 of the element itself, barrier, phase 2 (threads outside the domain take part with the tile's first
    // element and publish zeros, so that every thread reaches the barrier)
    ThIdx<th::dom_uw> idx;
    const bool inside = idx.from_coords(x0 + tx, y0 + ty, z0 + tz);
    if (!inside) idx.from_coords(x0, y0, z0);
    TAcc<th::dom_uw, UPD, EDGE> a(idx, sm, inside ? tx : 0, inside ? ty : 0, inside ? tz : 0, beta);
    ThJp J{jpbuf, tx, ty, tz};
    real o[TH_U];
    th::applyJTJ_tile(a, P, J, inside, o);
    real dot = (real)0;
    if (!inside) return dot;
#if TH_MULTI
    if (!th_owned_slow(idx.c[th::dom_uw::ND - 1])) {       // ghost layer: keep the local copy of the search direction current
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < TH_NUM_UIMG; ++k) {
                const int te = a.template tile_elem<0, 0, 0>(TH_VTILE[k].roww, TH_VTILE[k].padl, TH_UIMG[k].channels);
#pragma unroll
                for (int ch = 0; ch < TH_UIMG[k].channels; ++ch)
                    pnew[TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch] = a.vec_at(k, te + ch);
            }
        }
        return dot;
    }
#endif
    if (!th::exclude_u0(a, P)) {
        int j = 0;
#pragma unroll
        for (int k = 0; k < TH_NUM_UIMG; ++k) {
            const int te = a.template tile_elem<0, 0, 0>(TH_VTILE[k].roww, TH_VTILE[k].padl, TH_UIMG[k].channels);
#pragma unroll
            for (int ch = 0; ch < TH_UIMG[k].channels; ++ch, ++j) {
                const long long off = TH_UIMG[k].offset + idx.lin * TH_UIMG[k].channels + ch;
                const real pv = a.vec_at(k, te + ch);
                real val = o[j];
#if TH_LM
#if TH_STAGE_CTC
                val += ((const real*)(sm + TH_VTILE[k].coff))[a.center_elem(TH_VTILE[k].croww, TH_UIMG[k].channels) + ch] * pv;
#else
                val += V.CtC[off] * pv;
#endif
#endif
                out[off] = val;
                if (mode == 0) pnew[off] = pv;
                dot += pv * val;
            }
        }
    }
    return dot;
}
#endif

template <bool TMA>
__device__ __forceinline__ void th_pcg_a_impl(const Params& P, const Vecs& V, const ThMaps& M, ThScalars* S, double* partials, int mode,
                                              const ThPeers& R) {
    extern __shared__ __align__(128) unsigned char th_sm[];      // TMA: TH_PIPE stages of TH_SMEM_BYTES; otherwise one
    __shared__ __align__(8) unsigned long long full[TH_PIPE], empty[TH_PIPE];
    if (S->done) return;
    const int it = S->it;
    const int tx = threadIdx.x, ty = threadIdx.y, tz = threadIdx.z;
    const int tid = tx + TH_TW * (ty + TH_TH * tz);
    const bool upd = mode == 0 && it > 0;
    const real beta = upd ? th_beta_prev(S) : (real)0;
    real* __restrict__ out = mode ? V.Adelta : V.Ap;
    real* __restrict__ pnew = ((it + 1) & 1) ? V.p2 : V.p;
    if (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < TH_PIPE; ++s) { th_mbar_init(&full[s], 1); th_mbar_init(&empty[s], TH_TILE_THREADS / 32); }
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < TH_PIPE; ++s) {
                const int t0 = (int)blockIdx.x + s * (int)gridDim.x;
                if (t0 < TH_NTILES) th_tile_issue(th_sm + s * TH_SMEM_BYTES, &full[s], M, t0, mode, it);
            }
        }
    }
    double acc[1] = {0.0};
    int tile_k = 0;
    (void)tile_k;
    unsigned phase = 0;                   // bit s: parity of the next completion of full[s] (and of empty[s])
    int stage = 0;
    for (int t = blockIdx.x; t < TH_NTILES; t += gridDim.x) {
        unsigned char* sm = th_sm + (TMA ? stage * TH_SMEM_BYTES : 0);
        if (TMA) {
            if (TH_PIPE == 2) {
                // two stages: the other stage held the previous tile; once every warp has released
                // it, refill it with the next tile so that the copy overlaps this tile's arithmetic
                const int tn = t + 2 * (int)gridDim.x, so = stage ^ 1;
                if (tid == 0 && t != (int)blockIdx.x && tn - (int)gridDim.x < TH_NTILES) {
                    th_mbar_wait(&empty[so], ((phase >> so) & 1u) ^ 1u);
                    th_tile_issue(th_sm + so * TH_SMEM_BYTES, &full[so], M, tn - (int)gridDim.x, mode, it);
                }
            }
            th_mbar_wait(&full[stage], (phase >> stage) & 1u);
        } else {
            __syncthreads();                         // previous tile fully consumed
            th_tile_fill(sm, P, V, t, mode, it, tid);
            __syncthreads();
        }
        const bool edge = th_tile_is_edge(t);          // uniform over the CTA
        real dot;
#if TH_TWO_PHASE
        real* jpbuf = (real*)(th_sm + (TMA ? TH_PIPE : 1) * TH_SMEM_BYTES) + (TH_JP_BUFS == 2 ? (tile_k & 1) * (TH_JP_NT * TH_JP_NBOX) : 0);
        ++tile_k;
        if (edge) dot = upd ? th_tile_apply2<true, true>(sm, jpbuf, P, V, t, mode, beta, out, pnew, tx, ty, tz)
                            : th_tile_apply2<false, true>(sm, jpbuf, P, V, t, mode, beta, out, pnew, tx, ty, tz);
        else dot = upd ? th_tile_apply2<true, false>(sm, jpbuf, P, V, t, mode, beta, out, pnew, tx, ty, tz)
                       : th_tile_apply2<false, false>(sm, jpbuf, P, V, t, mode, beta, out, pnew, tx, ty, tz);
#else
        if (edge) dot = upd ? th_tile_apply<true, true>(sm, P, V, t, mode, beta, out, pnew, tx, ty, tz)
                            : th_tile_apply<false, true>(sm, P, V, t, mode, beta, out, pnew, tx, ty, tz);
        else dot = upd ? th_tile_apply<true, false>(sm, P, V, t, mode, beta, out, pnew, tx, ty, tz)
                       : th_tile_apply<false, false>(sm, P, V, t, mode, beta, out, pnew, tx, ty, tz);
#endif
        acc[0] += (double)dot;
        if (TMA) {
            __syncwarp();
            if ((tid & 31) == 0) th_mbar_arrive(&empty[stage]);
            if (TH_PIPE == 1) {
                // one stage (tiles too large for two): refill as soon as every warp has released it;
                // the other CTAs resident on the SM cover the copy
                const int tn = t + (int)gridDim.x;
                if (tid == 0 && tn < TH_NTILES) {
                    th_mbar_wait(&empty[0], phase & 1u);
                    th_tile_issue(sm, &full[0], M, tn, mode, it);
                }
            }
            phase ^= 1u << stage;
            stage = (stage + 1 == TH_PIPE) ? 0 : stage + 1;
        }
    }
    if (mode) return;
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[1])) th_ad_publish(S, tot, R, false, true);
}
extern "C" __global__ void __launch_bounds__(TH_TILE_THREADS, TH_PCG_A_MINB)
th_pcg_a(const __grid_constant__ Params P, const __grid_constant__ Vecs V, const __grid_constant__ ThMaps M, ThScalars* S, double* partials, int mode,
         const __grid_constant__ ThPeers R) {
    th_pcg_a_impl<true>(P, V, M, S, partials, mode, R);
}
extern "C" __global__ void __launch_bounds__(TH_TILE_THREADS, TH_PCG_A_MINB)
th_pcg_a_ld(const __grid_constant__ Params P, const __grid_constant__ Vecs V, const __grid_constant__ ThMaps M, ThScalars* S, double* partials, int mode,
            const __grid_constant__ ThPeers R) {
    th_pcg_a_impl<false>(P, V, M, S, partials, mode, R);
}
#endif  // TH_TILED

// ------------------------------------------------------------------ per-unknown-image dispatch for flat kernels
template <int K> struct ThExclude;
#define TH_EXCL_CASE(K) \
    template <> struct ThExclude<K> { static __device__ __forceinline__ bool get(long long e, const Params& P) { \
        ThIdx<th::dom_u##K> i; i.from_linear(e); GAcc<th::dom_u##K> a(i, nullptr); return th::exclude_u##K(a, P); } };
TH_EXCL_CASE(0)
#if TH_NUM_UIMG > 1
TH_EXCL_CASE(1)
#endif
#if TH_NUM_UIMG > 2
TH_EXCL_CASE(2)
#endif
#if TH_NUM_UIMG > 3
TH_EXCL_CASE(3)
#endif
#if TH_NUM_UIMG > 4
#error "more than 4 unknown images: extend TH_EXCL_CASE"
#endif

template <int K> __device__ __forceinline__ bool th_excluded_rec(long long f, const Params& P) {
    if (f < TH_UIMG[K].offset + TH_UIMG[K].elements * TH_UIMG[K].channels)
        return ThExclude<K>::get((f - TH_UIMG[K].offset) / TH_UIMG[K].channels, P);
    if (K + 1 < TH_NUM_UIMG) return th_excluded_rec<(K + 1 < TH_NUM_UIMG ? K + 1 : K)>(f, P);
    return false;
}
__device__ __forceinline__ bool th_excluded(long long f, const Params& P) { return th_excluded_rec<0>(f, P); }
template <int K> __device__ __forceinline__ real* th_xptr_rec(long long f, const Params& P) {
    if (f < TH_UIMG[K].offset + TH_UIMG[K].elements * TH_UIMG[K].channels)
        return ((real*)P.ptr[TH_UIMG[K].ptr_slot]) + (f - TH_UIMG[K].offset);
    if (K + 1 < TH_NUM_UIMG) return th_xptr_rec<(K + 1 < TH_NUM_UIMG ? K + 1 : K)>(f, P);
    return nullptr;
}

// X += delta, honouring exclude so that excluded caller-owned unknowns are never written
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_update(const __grid_constant__ Params P, const __grid_constant__ Vecs V) {
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < TH_NUNK; f += (long long)gridDim.x * blockDim.x) {
        if (th_excluded(f, P)) continue;
        real* x = th_xptr_rec<0>(f, P);
        *x = *x + V.delta[f];
    }
}
// dir 0: buf = X (savePreviousUnknowns / initX); dir 1: X = buf (revertUpdate / reset_unknowns)
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_copy_x(const __grid_constant__ Params P, real* buf, int dir) {
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < TH_NUNK; f += (long long)gridDim.x * blockDim.x) {
        if (th_excluded(f, P)) continue;
        real* x = th_xptr_rec<0>(f, P);
        if (dir) *x = buf[f]; else buf[f] = *x;
    }
}

// ================================================================== residualwise-schedule unknown passes
// After the scatter kernels: V.r holds -J^T F, V.pre holds diag(J^T J) (raw).
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_init_finish(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials, int first_nonlinear,
               const __grid_constant__ ThPeers R, const __grid_constant__ ThPush H) {
    double acc[1] = {0.0};
    bool pushed = false;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < TH_NUNK; f += (long long)gridDim.x * blockDim.x) {
        // ghost entries (graph partition) are the owner's to initialise, p arrives by its push; delta is maintained
        // locally (th_pcg_b) and restarts from zero
        if (!th_flat_owned(f)) { V.delta[f] = (real)0; continue; }
        if (th_excluded(f, P)) { th_zero_scalar(V, H, pushed, f); continue; }
        const real rp = th_init_scalar(P, V, H, pushed, f, -V.r[f], V.pre[f], (real)1, first_nonlinear);   // pre := 1 when off, gauss_newton.t:718-722
        if (th_flat_counted(f)) acc[0] += (double)rp;
    }
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[0], pushed)) th_init_publish(S, tot, R, H, TH_TILED ? V.z : V.p);
}

// which = 0: finish Ap (LM: += CtC p) and alphaDenominator; which = 1: only mask Adelta
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_step1_finish(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials, int which) {
    if (S->done) return;
    double acc[1] = {0.0};
    real* __restrict__ out = which ? V.Adelta : V.Ap;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < TH_NUNK; f += (long long)gridDim.x * blockDim.x) {
        if (th_excluded(f, P)) { out[f] = (real)0; continue; }
        if (which) continue;
        const real p = V.p[f];
        real ap = out[f];
#if TH_LM
        ap += V.CtC[f] * p;
        out[f] = ap;
#endif
        acc[0] += (double)(p * ap);
    }
    if (which) return;
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[1])) {
        if (threadIdx.x == 0) S->aD = tot[0];
    }
}

// ================================================================== gather schedule (graph domains, materialised J)
#if TH_GATHER
__device__ constexpr ThSpace TH_SPACE[TH_NSPACES] = TH_SPACE_TABLE;
__device__ constexpr ThSlot TH_SLOT[TH_NSPACES][TH_MAXSLOTS] = TH_SLOT_TABLE;
__device__ constexpr int TH_NNZP[TH_NGROUPS] = TH_GROUP_NNZP;

#if TH_HAS_REP
__device__ constexpr int TH_SPACE_REP_T[TH_NSPACES] = TH_SPACE_REP;
#define TH_SPACE_IS_REP(SP) (TH_SPACE_REP_T[SP] != 0)
#else
#define TH_SPACE_IS_REP(SP) false
#endif
// sum over the LANES adjacent lanes that share one unknown element (LANES a power of two)
template <int LANES> __device__ __forceinline__ real th_lanes_sum_real(real v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Operator of one index space: Ap[n] = sum over the residual elements incident to unknown element
// n of (their partials at n) * (J p of that residual) [+ CtC p in LM], and <p, Ap>.  Replaces the
// clear of Ap + PCGStep1 residualwise (gauss_newton.t:1006-1016) + PCGStep1_Finish (:774-799), and
// for materialised groups the second (transposed) csrmv of cusparseJTJMatVec (:1497-1510).
// TH_SPACE[].lanes adjacent lanes share one unknown element and split its adjacency list between
// them: 1 for low-degree spaces whose walk is compute-bound, 2-8 to shorten the chain of dependent
// loads (index -> neighbour data) each thread walks, 32 (a warp) when many residuals meet at an
// element (the cameras of bundle adjustment).  which = 1: Adelta = A delta.
#define TH_GATHER_KERNEL(SP)                                                                                        \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_gather_s##SP(const __grid_constant__ Params P, const __grid_constant__ Vecs V,                               \
                    const __grid_constant__ ThGather G, ThScalars* S, double* partials, int which, int first,       \
                    int reduce_now, const __grid_constant__ ThPeers R) {                                            \
        if (S->done) return;                                                                                        \
        constexpr int LANES = TH_SPACE[SP].lanes;                                                                   \
        constexpr int NS = TH_SPACE[SP].nslots;                                                                     \
        const int lane = (int)(threadIdx.x % LANES);                                                                \
        const real* __restrict__ in = which ? V.delta : V.p;                                                        \
        real* __restrict__ out = which ? V.Adelta : V.Ap;                                                           \
        double acc1[1] = {0.0};                                                                                     \
        /* persistent blocks (grid = SMs x resident blocks): one grid reduction per block instead of one per 256 */ \
        /* elements; whole warps stay in the loop together (the lane sums below shuffle)                         */ \
        const long long total = (TH_SPACE[SP].elements * LANES + 31) / 32 * 32;                                     \
        for (long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x; gt < total;                           \
             gt += (long long)gridDim.x * blockDim.x) {                                                             \
            ThIdx<th::dom_s##SP> t;                                                                                 \
            const bool valid = t.from_linear(gt / LANES) && th_owned(t);   /* ghost vertices: the owner computes Ap */ \
            bool ex = true;                                                                                         \
            real acc[NS];                                                                                           \
            _Pragma("unroll") for (int j = 0; j < NS; ++j) acc[j] = (real)0;                                        \
            if (valid) {                                                                                            \
                GAcc<th::dom_s##SP> ta(t, nullptr);                                                                 \
                ex = th::exclude_s##SP(ta, P);                                                                      \
                if (!ex) {                                                                                          \
                    if (which) th::gather_s##SP<1, LANES>(t, lane, P, G, in, acc);                                  \
                    else th::gather_s##SP<0, LANES>(t, lane, P, G, in, acc);                                        \
                }                                                                                                   \
            }                                                                                                       \
            if (LANES > 1) { _Pragma("unroll") for (int j = 0; j < NS; ++j) acc[j] = th_lanes_sum_real<LANES>(acc[j]); } \
            if (valid && lane == 0) {                                                                               \
                real dot = (real)0;                                                                                 \
                _Pragma("unroll") for (int j = 0; j < NS; ++j) {                                                    \
                    const int k = TH_SLOT[SP][j].image;                                                             \
                    const long long off = TH_UIMG[k].offset + t.lin * TH_UIMG[k].channels + TH_SLOT[SP][j].channel; \
                    if (ex) { out[off] = (real)0; continue; }                                                       \
                    if (TH_SPACE_IS_REP(SP)) { out[off] = acc[j]; continue; }   /* summed over the ranks, then th_rep_finish */ \
                    const real pv = in[off];                                                                        \
                    real val = acc[j];                                                                              \
                    if (TH_LM) val += V.CtC[off] * pv;                                                              \
                    out[off] = val;                                                                                 \
                    dot += pv * val;                                                                                \
                }                                                                                                   \
                acc1[0] += (double)dot;                                                                             \
            }                                                                                                       \
        }                                                                                                           \
        if (which) return;                                                                                          \
        double tot[1];                                                                                              \
        if (th_grid_reduce<1>(acc1, tot, partials, &S->ticket[1])) th_ad_publish(S, tot, R, !first, reduce_now != 0); \
    }
TH_SPACE_LIST(TH_GATHER_KERNEL)

// PCGInit1 of the gather schedule: r = -J^T F and the raw diagonal of J^T J, gathered per unknown element over the
// same adjacency lists (replaces the clears of r and the preconditioner + the scattering residualwise PCGInit1,
// gauss_newton.t:998-1004; no atomics, deterministic).  th_init_finish then proceeds as after the scatter form.
#define TH_GATHERJTF_KERNEL(SP)                                                                                     \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_gatherjtf_s##SP(const __grid_constant__ Params P, const __grid_constant__ Vecs V,                            \
                       const __grid_constant__ ThGather G) {                                                        \
        constexpr int LANES = TH_SPACE[SP].lanes;                                                                   \
        constexpr int NS = TH_SPACE[SP].nslots;                                                                     \
        const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;                                      \
        const int lane = (int)(threadIdx.x % LANES);                                                                \
        ThIdx<th::dom_s##SP> t;                                                                                     \
        const bool valid = t.from_linear(gt / LANES) && th_owned(t);                                                \
        real accg[NS], accd[NS];                                                                                    \
        _Pragma("unroll") for (int j = 0; j < NS; ++j) { accg[j] = (real)0; accd[j] = (real)0; }                    \
        if (valid) th::gatherjtf_s##SP<LANES>(t, lane, P, G, accg, accd);                                           \
        if (LANES > 1) {                                                                                            \
            _Pragma("unroll") for (int j = 0; j < NS; ++j) {                                                        \
                accg[j] = th_lanes_sum_real<LANES>(accg[j]); accd[j] = th_lanes_sum_real<LANES>(accd[j]);           \
            }                                                                                                       \
        }                                                                                                           \
        if (valid && lane == 0) {                                                                                   \
            _Pragma("unroll") for (int j = 0; j < NS; ++j) {                                                        \
                const int k = TH_SLOT[SP][j].image;                                                                 \
                const long long off = TH_UIMG[k].offset + t.lin * TH_UIMG[k].channels + TH_SLOT[SP][j].channel;     \
                V.r[off] = accg[j];                                                                                 \
                V.pre[off] = accd[j];                                                                               \
            }                                                                                                       \
        }                                                                                                           \
    }
TH_SPACE_LIST(TH_GATHERJTF_KERNEL)

#if TH_HAS_REP
// Replicated unknowns (the cameras of a point-partitioned bundle adjustment): the gather kernels left this rank's
// partial sums in Ap / Adelta, NCCL summed them over the ranks; finish like the gather kernel does for
// partitioned unknowns: + CtC p in LM and, on the one rank that counts them, their part of <p, Ap>.
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_rep_finish(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials, int which,
              const __grid_constant__ ThPeers R) {
    if (S->done) return;
    const real* __restrict__ in = which ? V.delta : V.p;
    real* __restrict__ out = which ? V.Adelta : V.Ap;
    double acc1[1] = {0.0};
#pragma unroll
    for (int k = 0; k < TH_NUM_UIMG; ++k) {
        if (!TH_REP[k]) continue;
        const long long lo = TH_UIMG[k].offset, hi = lo + TH_UIMG[k].elements * TH_UIMG[k].channels;
        for (long long f = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; f < hi; f += (long long)gridDim.x * blockDim.x) {
            if (th_excluded(f, P)) { out[f] = (real)0; continue; }
            const real pv = in[f];
            real val = out[f];
            if (TH_LM) val += V.CtC[f] * pv;
            out[f] = val;
            if (TH_REP_OWNER) acc1[0] += (double)(pv * val);
        }
    }
    if (which) return;
    double tot[1];
    if (th_grid_reduce<1>(acc1, tot, partials, &S->ticket[1])) th_ad_publish(S, tot, R, true, true);
}
#endif

// Materialised Jacobian of one residual group (sparse_materialize schedules): the partial
// derivatives are stored once per nonlinear iteration (precomputeJ, gauss_newton.t:1019-1025; no
// sort / transpose as in cusparseOuter :1332-1446 -- the column structure is the caller's index
// arrays), and every PCG iteration forms J p per residual row from the stored values (the first
// csrmv of cusparseJTJMatVec, :1478-1490).
#define TH_MAT_KERNELS(G_)                                                                                          \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_computejv_g##G_(const __grid_constant__ Params P, const __grid_constant__ ThGather G) {                      \
        ThIdx<th::dom_g##G_> idx;                                                                                   \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G_> a(idx, nullptr);                                                                    \
            real jv[TH_NNZP[G_]];                                                                                   \
            th::computeJv_g##G_(a, P, jv);                                                                          \
            real4* __restrict__ dst = (real4*)(G.jvals[G_] + idx.lin * TH_NNZP[G_]);                                \
            _Pragma("unroll") for (int i = 0; i < TH_NNZP[G_] / 4; ++i) {                                           \
                real4 q; q.x = jv[4 * i]; q.y = jv[4 * i + 1]; q.z = jv[4 * i + 2]; q.w = jv[4 * i + 3];            \
                dst[i] = q;                                                                                         \
            }                                                                                                       \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_matj_g##G_(const __grid_constant__ Params P, const __grid_constant__ Vecs V,                                 \
                  const __grid_constant__ ThGather G, const ThScalars* S) {                                         \
        if (S->done) return;                                                                                        \
        ThIdx<th::dom_g##G_> idx;                                                                                   \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G_> a(idx, V.p);                                                                        \
            real jpv[TH_GROUPS[G_].nterms];                                                                         \
            th::matJ_g##G_(a, P, G.jvals[G_] + idx.lin * TH_NNZP[G_], jpv);                                         \
            _Pragma("unroll") for (int t = 0; t < TH_GROUPS[G_].nterms; ++t)                                        \
                G.jp[G_][idx.lin * TH_GROUPS[G_].nterms + t] = jpv[t];                                              \
        }                                                                                                           \
    }
TH_MAT_LIST(TH_MAT_KERNELS)

// Jt[Jp] schedule of one residual group (APPLY_SEPARATELY, thallo.t:4121): J p per residual row, matrix-free,
// stored for the gather kernels, which apply the transposed partials.  PCGStep1_J (gauss_newton.t:1027-1034);
// the scattering PCGStep1_Jt (:1036-1047) and the clear of Jp (:1646) are not needed.
#ifdef TH_JP_LIST
#define TH_JP_KERNEL(G_)                                                                                            \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_applyj_g##G_(const __grid_constant__ Params P, const __grid_constant__ Vecs V,                               \
                    const __grid_constant__ ThGather G, const ThScalars* S) {                                       \
        if (S->done) return;                                                                                        \
        ThIdx<th::dom_g##G_> idx;                                                                                   \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G_> a(idx, V.p);                                                                        \
            real jpv[TH_GROUPS[G_].nterms];                                                                         \
            th::applyJ_g##G_(a, P, jpv);                                                                            \
            _Pragma("unroll") for (int t = 0; t < TH_GROUPS[G_].nterms; ++t)                                        \
                G.jp[G_][idx.lin * TH_GROUPS[G_].nterms + t] = jpv[t];                                              \
        }                                                                                                           \
    }
TH_JP_LIST(TH_JP_KERNEL)
#endif

// Hoisted per-element invariants of one index space (transcendentals of a single unknown element, e.g.
// sin/cos of a vertex's angles): evaluated once per nonlinear iteration into the plan-owned image
// __coef_s<i>, which the endpoint functions read through their own index.
__device__ constexpr int TH_SCOEF_CH[TH_NSPACES] = TH_SCOEF_N;
__device__ constexpr int TH_SCOEF_PTR[TH_NSPACES] = TH_SCOEF_SLOT;
#define TH_SCOEF_KERNEL(SP)                                                                                         \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_precompute_scoef_s##SP(const __grid_constant__ Params P) {                                                   \
        ThIdx<th::dom_s##SP> idx;                                                                                   \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_s##SP> a(idx, nullptr);                                                                    \
            real c[TH_SCOEF_CH[SP]];                                                                                \
            th::scoef_s##SP(a, P, c);                                                                               \
            real* __restrict__ out = (real*)P.ptr[TH_SCOEF_PTR[SP]];                                                \
            _Pragma("unroll") for (int i = 0; i < TH_SCOEF_CH[SP]; ++i) out[idx.lin * TH_SCOEF_CH[SP] + i] = c[i];  \
        }                                                                                                           \
    }
TH_SCOEF_LIST(TH_SCOEF_KERNEL)

// 64-bit checksum of an index array (sum of a mixing hash of every (position, value) pair): lets the plan notice that a caller changed
// the contents of a sparse index array between solves (the adjacency lists are then rebuilt)
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_index_checksum(const int* __restrict__ a, long long n, unsigned long long* out) {
    unsigned long long h = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {   // splitmix64 of (position, value): no linear relation between entries survives (a_1 += 5, a_7 -= 1 changed nothing before)
        unsigned long long x = ((unsigned long long)i << 32) ^ (unsigned long long)(unsigned int)__ldg(a + i);
        x += 0x9E3779B97F4A7C15ull;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        h += x ^ (x >> 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}
#endif  // TH_GATHER

// ================================================================== computed arrays
// precompute (gauss_newton.t:979-986, createprecomputed thallo.t:4046-4094): every ComputedArray
// (`exp:get(...)`) is stored as a plan-owned image together with its gradient image, refreshed
// whenever the unknowns change (init, after every update, after a rejected LM step).
#if TH_NCOMPUTED > 0
struct ThComputed { int val_slot; int grad_slot; int ngrad; };
__device__ constexpr ThComputed TH_COMPUTED[TH_NCOMPUTED] = TH_COMPUTED_TABLE;
#define TH_COMPUTED_KERNEL(K)                                                                                       \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_precompute_c##K(const __grid_constant__ Params P) {                                                          \
        ThIdx<th::dom_c##K> idx;                                                                                    \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_c##K> a(idx, nullptr);                                                                     \
            real v[1 + TH_COMPUTED[K].ngrad];                                                                       \
            th::precompute_c##K(a, P, v);                                                                           \
            ((real*)P.ptr[TH_COMPUTED[K].val_slot])[idx.lin] = v[0];                                                \
            if (TH_COMPUTED[K].ngrad > 0) {                                                                         \
                real* __restrict__ g = (real*)P.ptr[TH_COMPUTED[K].grad_slot >= 0 ? TH_COMPUTED[K].grad_slot : 0];  \
                _Pragma("unroll") for (int i = 0; i < TH_COMPUTED[K].ngrad; ++i)                                    \
                    g[idx.lin * TH_COMPUTED[K].ngrad + i] = v[1 + i];                                               \
            }                                                                                                       \
        }                                                                                                           \
    }
TH_COMPUTED_LIST(TH_COMPUTED_KERNEL)
#endif

// ================================================================== per-residual-group kernels
#define TH_GROUP_KERNELS(G)                                                                                         \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_cost_g##G(const __grid_constant__ Params P, ThScalars* S, double* partials, int first) {                     \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        double acc[1] = {0.0};                                                                                      \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x) &&                                    \
            th_owned(idx)) {                                                                                        \
            GAcc<th::dom_g##G> a(idx, nullptr);                                                                     \
            acc[0] = (double)th::cost_g##G(a, P);                                                                   \
        }                                                                                                           \
        double tot[1];                                                                                              \
        if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[4])) {                                                 \
            if (threadIdx.x == 0) S->cost = (first ? 0.0 : S->cost) + tot[0];                                       \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_modelcost_g##G(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S,               \
                      double* partials, int first) {                                                                \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        double acc[1] = {0.0};                                                                                      \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x) &&                                    \
            th_owned(idx)) {                                                                                        \
            GAcc<th::dom_g##G> a(idx, V.delta);                                                                     \
            acc[0] = (double)th::modelcost_g##G(a, P);                                                              \
        }                                                                                                           \
        double tot[1];                                                                                              \
        if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[4])) {                                                 \
            if (threadIdx.x == 0) S->modelcost = (first ? 0.0 : S->modelcost) + tot[0];                             \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_evaljtf_g##G(const __grid_constant__ Params P, const __grid_constant__ Vecs V) {                             \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G> a(idx, nullptr);                                                                     \
            GScatter<th::dom_g##G> s(idx, V.r, V.pre);                                                              \
            th::evalJTF_g##G(a, P, s);                                                                              \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_applyjtj_g##G(const __grid_constant__ Params P, const __grid_constant__ Vecs V, const ThScalars* S, int which) { \
        if (S->done) return;                                                                                        \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G> a(idx, which ? V.delta : V.p);                                                       \
            GScatter<th::dom_g##G> s(idx, which ? V.Adelta : V.Ap, nullptr);                                        \
            th::applyJTJ_g##G(a, P, s);                                                                             \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_residuals_g##G(const __grid_constant__ Params P, real* out) {                                                \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G> a(idx, nullptr);                                                                     \
            real r[TH_GROUPS[G].nterms];                                                                            \
            th::residuals_g##G(a, P, r);                                                                            \
            for (int t = 0; t < TH_GROUPS[G].nterms; ++t) out[idx.lin * TH_GROUPS[G].nterms + t] = r[t];            \
        }                                                                                                           \
    }                                                                                                               \
    extern "C" __global__ void __launch_bounds__(TH_BLOCK)                                                          \
    th_computej_g##G(const __grid_constant__ Params P, real* vals, long long* cols) {                               \
        ThIdx<th::dom_g##G> idx;                                                                                    \
        if (idx.from_linear((long long)blockIdx.x * blockDim.x + threadIdx.x)) {                                    \
            GAcc<th::dom_g##G> a(idx, nullptr);                                                                     \
            real v[TH_GROUPS[G].nnz > 0 ? TH_GROUPS[G].nnz : 1];                                                    \
            long long c[TH_GROUPS[G].nnz > 0 ? TH_GROUPS[G].nnz : 1];                                               \
            th::computeJ_g##G(a, P, v, c);                                                                          \
            for (int t = 0; t < TH_GROUPS[G].nnz; ++t) {                                                            \
                vals[idx.lin * TH_GROUPS[G].nnz + t] = v[t];                                                        \
                cols[idx.lin * TH_GROUPS[G].nnz + t] = c[t];                                                        \
            }                                                                                                       \
        }                                                                                                           \
    }
TH_GROUP_LIST(TH_GROUP_KERNELS)

#if TH_AT_OUTPUT
// At-output plans: every residual group lives on the unknown domain, so the cost and the model cost of ALL groups
// are formed in one pass (the unknowns, delta and the auxiliary images are read once instead of once per group;
// image_warping: 5 + 5 launches per nonlinear iteration -> 1 + 1).  Per element the groups' values are the same
// numbers as in th_cost_g<i> / th_modelcost_g<i>, added up in double.
#define TH_COST_TERM(G) + (double)th::cost_g##G(a, P)
#define TH_MODELCOST_TERM(G) + (double)th::modelcost_g##G(a, P)
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_cost_uw(const __grid_constant__ Params P, ThScalars* S, double* partials) {
    ThIdx<th::dom_uw> idx;
    double acc[1] = {0.0};
    if (th_uw_index(idx) && th_owned(idx)) {
        GAcc<th::dom_uw> a(idx, nullptr);
        acc[0] = 0.0 TH_GROUP_LIST(TH_COST_TERM);
    }
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[4])) {
        if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) S->cost = tot[0];
    }
}
extern "C" __global__ void __launch_bounds__(TH_BLOCK)
th_modelcost_uw(const __grid_constant__ Params P, const __grid_constant__ Vecs V, ThScalars* S, double* partials) {
    ThIdx<th::dom_uw> idx;
    double acc[1] = {0.0};
    if (th_uw_index(idx) && th_owned(idx)) {
        GAcc<th::dom_uw> a(idx, V.delta);
        acc[0] = 0.0 TH_GROUP_LIST(TH_MODELCOST_TERM);
    }
    double tot[1];
    if (th_grid_reduce<1>(acc, tot, partials, &S->ticket[4])) {
        if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) S->modelcost = tot[0];
    }
}
#endif
