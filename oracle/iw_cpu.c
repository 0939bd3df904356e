/* ORACLE (test infrastructure and CPU baseline, not product code).
 *
 * Plain-C restatement of the reference's CPU path (cpuOnly=1) for the image_warping energy
 * (config 2 of BASELINE.json): the GN / LM outer loop and the PCG inner loop of
 * reference API/src/gauss_newton.t executed the way API/src/cpu_cuda.t:265-301 executes
 * "kernels" -- nested loops over blocks (x fastest) and the threads of each block
 * (x fastest; unknownwise 2-D kernels use 16x16 blocks, util.t:715-725), atomics as plain
 * `+=` (cpu_cuda.t:110-124), warp reductions as identity (cpu_cuda.t:114-120), i.e. every
 * dot product is one running sum in tile order.
 *
 * What each function follows (reference file:line):
 *   energy                examples/image_warping/image_warping.t:1-34
 *   iw_cost               computeCost gauss_newton.t:1067-1079 + createcost thallo.t:3939-3949
 *   eval_jtf              evalJTFUnknownwise thallo.t:3669-3712  (sum d.r, sum d^2 at the unknown)
 *   apply_jtj             applyJTJUnknownwise thallo.t:3603-3667 (J^T J p gathered at the unknown)
 *   pcg_init              PCGInit1 unknownwise gauss_newton.t:678-710, guardedInvert CERES :641-648,
 *                         LM: PCGSaveSSq :929-934, computeCtC thallo.t:3911-3937, PCGFinalizeDiagonal :936-969
 *   pcg_step1/2/3         PCGStep1 :734-752, PCGStep2 :801-843 (reset variant :845-886), PCGStep3 :889-899
 *   model cost            createmodelcostResidualwise thallo.t:3845-3865, computeModelCost :1088-1095
 *   outer loop            step :1545-1785 (accept/reject :1707-1753), init :1166-1198
 * The per-pixel derivatives were derived by hand from the energy and are cross-checked
 * against the independent NumPy dual-number oracle (oracle/npdsl.py + oracle/solver.py,
 * itself pinned on the reference's golden PNGs) in tests/test_oracle_c.py.
 *
 * ACC is the accumulator type of the dot products: float reproduces the reference CPU
 * path's arithmetic (default, used for the timed baseline); -DACC=double is used by the
 * tests to compare trajectories with the NumPy oracle at 1e-5.
 * With OpenMP (-fopenmp, OMP_NUM_THREADS>1) the block loop is split across threads; this
 * is NOT the reference's behaviour (it is single-threaded by construction) and is only
 * used by `bench.py --impl reference` to give the CPU arm every host core.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ACC
#define ACC float
#endif
typedef float real;

typedef struct {
    int W, H;
    real* Offset;            /* W*H*2, unknown, updated in place */
    real* Angle;             /* W*H,   unknown, updated in place */
    const real* UrShape;     /* W*H*2 */
    const real* Constraints; /* W*H*2 */
    const real* Mask;        /* W*H   */
    real w_fitSqrt, w_regSqrt;
} IWProblem;

typedef struct {             /* gauss_newton.t:200-216, defaults :41-55 */
    int lm;                  /* 0 gauss_newton, 1 levenberg_marquardt (as written) */
    int nIterations, lIterations, residual_reset_period;
    real min_relative_decrease, min_trust_region_radius, max_trust_region_radius, q_tolerance,
        function_tolerance, trust_region_radius, radius_decrease_factor, min_lm_diagonal, max_lm_diagonal;
    long long max_pcg_iterations;   /* >0: stop the whole solve after this many PCG iterations (bounded sample) */
} IWSolverParams;

typedef struct {
    int n_nonlinear;           /* nonlinear iterations executed */
    long long n_pcg;           /* PCG iterations executed in total */
    double seconds_total;      /* wall clock of the whole solve */
    double seconds_pcg;        /* wall clock spent inside the PCG inner loops only */
    double cost[260];          /* cost[0] initial, cost[k] after nonlinear iteration k (LM) / final in cost[n_nonlinear+1] */
    int n_lin[256];            /* PCG iterations of nonlinear iteration k */
    int n_cost;
} IWResult;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

void iw_default_params(IWSolverParams* p) {
    p->lm = 0; p->nIterations = 10; p->lIterations = 10; p->residual_reset_period = 10;
    p->min_relative_decrease = 1e-3f; p->min_trust_region_radius = 1e-32f; p->max_trust_region_radius = 1e16f;
    p->q_tolerance = 0.0001f; p->function_tolerance = 0.000001f; p->trust_region_radius = 1e4f;
    p->radius_decrease_factor = 2.0f; p->min_lm_diagonal = 1e-6f; p->max_lm_diagonal = 1e32f;
    p->max_pcg_iterations = 0;
}

/* ---- image access: out-of-bounds get returns 0 (thallo.t:876-882) */
#define IDX(x, y) ((long long)(x) + (long long)W * (y))
static inline int inb(int W, int H, int x, int y) { return x >= 0 && x < W && y >= 0 && y < H; }
static inline real mask_at(const IWProblem* P, int x, int y) {
    return inb(P->W, P->H, x, y) ? P->Mask[(long long)x + (long long)P->W * y] : 0.0f;
}
static const int DX[4] = {1, -1, 0, 0}, DY[4] = {0, 0, 1, -1};

/* validity of the regularisation residual at (x,y) towards direction k (image_warping.t:19-27) */
static inline int reg_valid(const IWProblem* P, int x, int y, int k) {
    const int nx = x + DX[k], ny = y + DY[k];
    return inb(P->W, P->H, nx, ny) && mask_at(P, x, y) == 0.0f && mask_at(P, nx, ny) == 0.0f;
}
static inline int fit_valid(const IWProblem* P, long long i) {
    return P->Constraints[2 * i] >= 0.0f && P->Constraints[2 * i + 1] >= 0.0f && P->Mask[i] == 0.0f;
}
/* e = w_reg * ((O(c) - O(n)) - R(A(c)) (U(c) - U(n))) and a = d e / d A(c) */
static inline void reg_terms(const IWProblem* P, const real* Off, const real* Ang, long long c, long long n, real e[2], real a[2]) {
    const real w = P->w_regSqrt;
    const real cs = cosf(Ang[c]), sn = sinf(Ang[c]);
    const real u0 = P->UrShape[2 * c] - P->UrShape[2 * n], u1 = P->UrShape[2 * c + 1] - P->UrShape[2 * n + 1];
    e[0] = w * ((Off[2 * c] - Off[2 * n]) - (cs * u0 + (-sn) * u1));
    e[1] = w * ((Off[2 * c + 1] - Off[2 * n + 1]) - (sn * u0 + cs * u1));
    a[0] = -w * ((-sn) * u0 - cs * u1);
    a[1] = -w * (cs * u0 - sn * u1);
}

/* ---- the "launch" order of an unknownwise 2-D kernel on the CPU path: 16x16 blocks, x fastest */
#define TILE 16
#define FOR_TILES_BEGIN(W, H)                                                     \
    {                                                                             \
        const int _bx = ((W) + TILE - 1) / TILE, _by = ((H) + TILE - 1) / TILE;   \
        for (int _b = _lo; _b < _hi; ++_b) {                                      \
            const int _by0 = (_b / _bx) * TILE, _bx0 = (_b % _bx) * TILE;         \
            (void)_by;                                                            \
            for (int _ty = 0; _ty < TILE; ++_ty)                                  \
                for (int _tx = 0; _tx < TILE; ++_tx) {                            \
                    const int x = _bx0 + _tx, y = _by0 + _ty;                     \
                    if (x >= (W) || y >= (H)) continue;
#define FOR_TILES_END }}}

static int n_tiles(int W, int H) { return ((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE); }

/* split [0, ntiles) over OpenMP threads; thread t gets [lo, hi) */
static void tile_range(int nt, int* lo, int* hi) {
#ifdef _OPENMP
    const int T = omp_get_num_threads(), t = omp_get_thread_num();
#else
    const int T = 1, t = 0;
#endif
    const long long per = (nt + T - 1) / T;
    long long a = per * t, b = a + per;
    if (a > nt) a = nt;
    if (b > nt) b = nt;
    *lo = (int)a; *hi = (int)b;
}

/* cost = sum over residual groups and elements of 1/2 r^2 */
double iw_cost(const IWProblem* P) {
    const int W = P->W, H = P->H;
    ACC total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC acc = 0;
        /* the residualwise cost kernel is a flat 1-D launch over the residual domain; the order of a
           single running sum over pixels is what matters, tile order is kept for uniformity */
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            for (int k = 0; k < 4; ++k) {
                if (!reg_valid(P, x, y, k)) continue;
                real e[2], a[2];
                reg_terms(P, P->Offset, P->Angle, c, IDX(x + DX[k], y + DY[k]), e, a);
                acc += (ACC)(0.5f * (e[0] * e[0] + e[1] * e[1]));
            }
            if (fit_valid(P, c)) {
                const real f0 = P->w_fitSqrt * (P->Offset[2 * c] - P->Constraints[2 * c]);
                const real f1 = P->w_fitSqrt * (P->Offset[2 * c + 1] - P->Constraints[2 * c + 1]);
                acc += (ACC)(0.5f * (f0 * f0 + f1 * f1));
            }
        FOR_TILES_END
        total += acc;
    }
    return (double)(real)total;
}

typedef struct {
    long long n;      /* pixels */
    real *delta, *r, *b, *Adelta, *z, *p, *Ap, *CtC, *pre, *SSq, *prevX;   /* 3 reals per pixel: Ox, Oy, A */
} IWVecs;

/* J^T F and diag(J^T J) gathered at unknown (x,y) */
static inline void eval_jtf(const IWProblem* P, int x, int y, real g[3], real d[3]) {
    const int W = P->W;
    const long long c = IDX(x, y);
    const real w = P->w_regSqrt;
    g[0] = g[1] = g[2] = 0; d[0] = d[1] = d[2] = 0;
    for (int k = 0; k < 4; ++k) {
        if (!reg_valid(P, x, y, k)) continue;     /* validity is symmetric in (c, n): both instances exist or none */
        const long long n = IDX(x + DX[k], y + DY[k]);
        real e[2], a[2];
        reg_terms(P, P->Offset, P->Angle, c, n, e, a);      /* instance at c towards n */
        g[0] += w * e[0]; g[1] += w * e[1]; g[2] += a[0] * e[0] + a[1] * e[1];
        d[0] += w * w; d[1] += w * w; d[2] += a[0] * a[0] + a[1] * a[1];
        reg_terms(P, P->Offset, P->Angle, n, c, e, a);      /* instance at n towards c touches O(c) with -w */
        g[0] += -w * e[0]; g[1] += -w * e[1];
        d[0] += w * w; d[1] += w * w;
    }
    if (fit_valid(P, c)) {
        const real wf = P->w_fitSqrt;
        g[0] += wf * (wf * (P->Offset[2 * c] - P->Constraints[2 * c]));
        g[1] += wf * (wf * (P->Offset[2 * c + 1] - P->Constraints[2 * c + 1]));
        d[0] += wf * wf; d[1] += wf * wf;
    }
}

/* (J^T J v) gathered at unknown (x,y); v is an unknown-shaped vector (3 reals per pixel) */
static inline void apply_jtj(const IWProblem* P, const real* v, int x, int y, real out[3]) {
    const int W = P->W;
    const long long c = IDX(x, y);
    const real w = P->w_regSqrt;
    out[0] = out[1] = out[2] = 0;
    for (int k = 0; k < 4; ++k) {
        if (!reg_valid(P, x, y, k)) continue;
        const long long n = IDX(x + DX[k], y + DY[k]);
        real e[2], a[2], jp[2];
        reg_terms(P, P->Offset, P->Angle, c, n, e, a);
        jp[0] = w * v[3 * c] + (-w) * v[3 * n] + a[0] * v[3 * c + 2];
        jp[1] = w * v[3 * c + 1] + (-w) * v[3 * n + 1] + a[1] * v[3 * c + 2];
        out[0] += w * jp[0]; out[1] += w * jp[1]; out[2] += a[0] * jp[0] + a[1] * jp[1];
        reg_terms(P, P->Offset, P->Angle, n, c, e, a);
        jp[0] = w * v[3 * n] + (-w) * v[3 * c] + a[0] * v[3 * n + 2];
        jp[1] = w * v[3 * n + 1] + (-w) * v[3 * c + 1] + a[1] * v[3 * n + 2];
        out[0] += -w * jp[0]; out[1] += -w * jp[1];
    }
    if (fit_valid(P, c)) {
        const real wf = P->w_fitSqrt;
        out[0] += wf * (wf * v[3 * c]); out[1] += wf * (wf * v[3 * c + 1]);
    }
}

/* Solver vectors use the reference's unknown layout: Offset image (2 per pixel) then Angle image
   (thallo.t:1102-1126).  For the CPU restatement they are kept interleaved 3-per-pixel, which is a
   pure relabelling (no arithmetic depends on the storage order). */

static real guarded_invert(real d) { const real s = 1.0f + sqrtf(d); return 1.0f / (s * s); }

static ACC pcg_init(const IWProblem* P, IWVecs* V, const IWSolverParams* sp, real radius, int first) {
    const int W = P->W, H = P->H;
    ACC total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC acc = 0;
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;            /* excluded unknown: never touched, contributes 0 */
            real g[3], d[3];
            eval_jtf(P, x, y, g, d);
            for (int j = 0; j < 3; ++j) {
                const long long o = 3 * c + j;
                const real r = -g[j];
                real pre = guarded_invert(d[j]);
                if (sp->lm) {
                    if (first) V->SSq[o] = pre;
                    const real ssq = V->SSq[o];
                    const real ctc_raw = d[j] / radius;
                    const real mult = (1.0f / ssq) / radius;
                    const real ctc = fminf(fmaxf(ctc_raw, sp->min_lm_diagonal * mult), sp->max_lm_diagonal * mult);
                    pre = 1.0f / (ctc + radius * ctc_raw);
                    V->CtC[o] = ctc;
                    V->b[o] = r;
                }
                const real p = pre * r;
                V->delta[o] = 0; V->r[o] = r; V->pre[o] = pre; V->p[o] = p;
                acc += (ACC)(r * p);
            }
        FOR_TILES_END
        total += acc;
    }
    return total;
}

/* out = (J^T J [+ CtC]) in ; returns <in, out> */
static ACC pcg_step1(const IWProblem* P, IWVecs* V, const real* in, real* outv, int lm) {
    const int W = P->W, H = P->H;
    ACC total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC acc = 0;
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;
            real o[3];
            apply_jtj(P, in, x, y, o);
            for (int j = 0; j < 3; ++j) {
                const long long q = 3 * c + j;
                real val = o[j];
                if (lm) val += V->CtC[q] * in[q];
                outv[q] = val;
                acc += (ACC)(in[q] * val);
            }
        FOR_TILES_END
        total += acc;
    }
    return total;
}

static void pcg_step2(const IWProblem* P, IWVecs* V, real alpha, int lm, ACC* betaN, ACC* qout) {
    const int W = P->W, H = P->H;
    ACC tb = 0, tq = 0;
#pragma omp parallel reduction(+ : tb, tq)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC ab = 0, aq = 0;
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;
            for (int j = 0; j < 3; ++j) {
                const long long o = 3 * c + j;
                const real delta = V->delta[o] + alpha * V->p[o];
                V->delta[o] = delta;
                const real r = V->r[o] - alpha * V->Ap[o];
                V->r[o] = r;
                const real z = V->pre[o] * r;
                V->z[o] = z;
                ab += (ACC)(z * r);
                if (lm) aq += (ACC)(0.5f * (delta * (r + V->b[o])));
            }
        FOR_TILES_END
        tb += ab; tq += aq;
    }
    *betaN = tb; *qout = tq;
}

/* LM residual reset (gauss_newton.t:845-886): delta += alpha p; r = b - A delta */
static void pcg_step2_reset(const IWProblem* P, IWVecs* V, real alpha, ACC* betaN, ACC* qout) {
    const int W = P->W, H = P->H;
#pragma omp parallel
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;
            for (int j = 0; j < 3; ++j) V->delta[3 * c + j] = V->delta[3 * c + j] + alpha * V->p[3 * c + j];
        FOR_TILES_END
    }
    pcg_step1(P, V, V->delta, V->Adelta, 1);
    ACC tb = 0, tq = 0;
#pragma omp parallel reduction(+ : tb, tq)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC ab = 0, aq = 0;
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;
            for (int j = 0; j < 3; ++j) {
                const long long o = 3 * c + j;
                const real r = V->b[o] - V->Adelta[o];
                V->r[o] = r;
                const real z = V->pre[o] * r;
                V->z[o] = z;
                ab += (ACC)(z * r);
                aq += (ACC)(0.5f * (V->delta[o] * (r + V->b[o])));
            }
        FOR_TILES_END
        tb += ab; tq += aq;
    }
    *betaN = tb; *qout = tq;
}

static void pcg_step3(const IWProblem* P, IWVecs* V, real beta) {
    const int W = P->W, H = P->H;
#pragma omp parallel
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            if (P->Mask[c] != 0.0f) continue;
            for (int j = 0; j < 3; ++j) V->p[3 * c + j] = V->z[3 * c + j] + beta * V->p[3 * c + j];
        FOR_TILES_END
    }
}

static ACC model_cost(const IWProblem* P, const IWVecs* V) {
    const int W = P->W, H = P->H;
    const real w = P->w_regSqrt;
    const real* dl = V->delta;
    ACC total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int _lo, _hi;
        tile_range(n_tiles(W, H), &_lo, &_hi);
        ACC acc = 0;
        FOR_TILES_BEGIN(W, H)
            const long long c = IDX(x, y);
            for (int k = 0; k < 4; ++k) {
                if (!reg_valid(P, x, y, k)) continue;
                const long long n = IDX(x + DX[k], y + DY[k]);
                real e[2], a[2];
                reg_terms(P, P->Offset, P->Angle, c, n, e, a);
                const real m0 = e[0] + (w * dl[3 * c] + (-w) * dl[3 * n] + a[0] * dl[3 * c + 2]);
                const real m1 = e[1] + (w * dl[3 * c + 1] + (-w) * dl[3 * n + 1] + a[1] * dl[3 * c + 2]);
                acc += (ACC)(0.5f * (m0 * m0 + m1 * m1));
            }
            if (fit_valid(P, c)) {
                const real wf = P->w_fitSqrt;
                const real m0 = wf * (P->Offset[2 * c] - P->Constraints[2 * c]) + wf * dl[3 * c];
                const real m1 = wf * (P->Offset[2 * c + 1] - P->Constraints[2 * c + 1]) + wf * dl[3 * c + 1];
                acc += (ACC)(0.5f * (m0 * m0 + m1 * m1));
            }
        FOR_TILES_END
        total += acc;
    }
    return total;
}

static void update_x(IWProblem* P, const IWVecs* V, int save_prev, int revert) {
    const long long n = V->n;
#pragma omp parallel for
    for (long long c = 0; c < n; ++c) {
        if (P->Mask[c] != 0.0f) continue;
        if (revert) {
            P->Offset[2 * c] = V->prevX[3 * c]; P->Offset[2 * c + 1] = V->prevX[3 * c + 1]; P->Angle[c] = V->prevX[3 * c + 2];
            continue;
        }
        if (save_prev) { V->prevX[3 * c] = P->Offset[2 * c]; V->prevX[3 * c + 1] = P->Offset[2 * c + 1]; V->prevX[3 * c + 2] = P->Angle[c]; }
        P->Offset[2 * c] += V->delta[3 * c];
        P->Offset[2 * c + 1] += V->delta[3 * c + 1];
        P->Angle[c] += V->delta[3 * c + 2];
    }
}

void iw_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int iw_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thallo_ProblemSolve for image_warping on the CPU.  Returns 0 on success. */
int iw_solve(IWProblem* P, const IWSolverParams* sp_in, IWResult* R) {
    IWSolverParams sp = *sp_in;
    const long long n = (long long)P->W * P->H;
    IWVecs V;
    V.n = n;
    real* block = (real*)calloc((size_t)n * 3 * 11, sizeof(real));
    if (!block) return 1;
    real** slots[11] = {&V.delta, &V.r, &V.b, &V.Adelta, &V.z, &V.p, &V.Ap, &V.CtC, &V.pre, &V.SSq, &V.prevX};
    for (int i = 0; i < 11; ++i) *slots[i] = block + (size_t)n * 3 * i;
    memset(R, 0, sizeof(*R));
    const double t0 = now_s();
    real radius = sp.trust_region_radius, decrease_factor = sp.radius_decrease_factor;
    real prev_cost = (real)iw_cost(P);
    R->cost[R->n_cost++] = prev_cost;
    int nIter = 0, stop = 0;
    while (!stop && nIter < sp.nIterations) {
        ACC aN = pcg_init(P, &V, &sp, radius, nIter == 0);
        real Q0 = 0;
        int nlin = 0;
        const double tp = now_s();
        for (int l = 0; l < sp.lIterations; ++l) {
            const ACC aD = pcg_step1(P, &V, V.p, V.Ap, sp.lm);
            const real num = (real)aN, den = (real)aD;
            const real alpha = (sp.lm || den != 0.0f) ? num / den : 0.0f;     /* safeDivideIfNotLM :226-234 */
            ACC bN, q;
            if (sp.lm && ((l + 1) % sp.residual_reset_period) == 0) pcg_step2_reset(P, &V, alpha, &bN, &q);
            else pcg_step2(P, &V, alpha, sp.lm, &bN, &q);
            const real beta = (sp.lm || num != 0.0f) ? (real)bN / num : 0.0f;
            pcg_step3(P, &V, beta);
            aN = bN;
            ++nlin; ++R->n_pcg;
            if (sp.lm) {                                                     /* zeta test :1666-1686 */
                const real Q1 = (real)q;
                if (!isfinite(Q1)) break;
                const real zeta = (real)(l + 1) * (Q1 - Q0) / Q1;
                if (!isfinite(zeta) || zeta < sp.q_tolerance) break;
                Q0 = Q1;
            }
            if (sp.max_pcg_iterations > 0 && R->n_pcg >= sp.max_pcg_iterations) { stop = 1; break; }
        }
        R->seconds_pcg += now_s() - tp;
        R->n_lin[nIter < 256 ? nIter : 255] = nlin;
        if (stop) { ++nIter; break; }
        if (sp.lm) {
            const real mc = (real)model_cost(P, &V);
            const real model_cost_change = prev_cost - mc;
            update_x(P, &V, 1, 0);
            const real new_cost = (real)iw_cost(P);
            const real cost_change = prev_cost - new_cost;
            const real rel = cost_change / model_cost_change;
            if (R->n_cost < 259) R->cost[R->n_cost++] = new_cost;
            if (cost_change >= 0 && rel > sp.min_relative_decrease) {
                if (cost_change <= prev_cost * sp.function_tolerance) { ++nIter; break; }
                const double tmp = 1.0 - pow(2.0 * (double)rel - 1.0, 3.0);
                radius = (real)((double)radius / fmax(1.0 / 3.0, tmp));
                radius = (real)fmin((double)radius, (double)sp.max_trust_region_radius);
                decrease_factor = 2.0f;
                prev_cost = new_cost;
            } else {
                update_x(P, &V, 0, 1);
                radius = radius / decrease_factor;
                decrease_factor = 2.0f * decrease_factor;
                if (radius < sp.min_trust_region_radius) { ++nIter; break; }
            }
        } else {
            update_x(P, &V, 0, 0);
        }
        ++nIter;
    }
    R->n_nonlinear = nIter;
    if (R->n_cost < 260) R->cost[R->n_cost++] = iw_cost(P);     /* finalize :1200-1212 */
    R->seconds_total = now_s() - t0;
    free(block);
    return 0;
}
