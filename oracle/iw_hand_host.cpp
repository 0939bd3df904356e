// ORACLE (test infrastructure): host-side driver around the reference's HAND-DERIVED image_warping equations
// (examples/image_warping/src/WarpingSolverEquations.h: evalFDevice :9-43, evalMinusJTFDevice :49-200,
// applyJTJDevice :207-349), compiled from the reference tree where it lies (`make -C oracle hand`, output in
// oracle/_ref/, nothing copied).  The header is CUDA device code; this file turns the few device-only constructs it
// touches into host equivalents and calls the three functions for every pixel, so that the NumPy oracle's cost,
// -J^T F and J^T J p (dual-number AD of the energy as Thallo defines it) can be checked on the CPU against derivatives
// the reference's authors wrote by hand (SURVEY 8c "secondary oracle").  Scaling between the two: the hand solver
// minimises sum w e^2, Thallo 1/2 sum (sqrt(w) e)^2, so cost, gradient and J^T J of the former are twice the latter's.
#include "hand_shims.h"

#include REF_STATE_HEADER          // WarpingSolverState.h first, as WarpingSolver.cu includes it (the utility header needs SolverInput)
#include REF_EQUATIONS_HEADER

extern "C" {

struct Problem {
    SolverInput in;
    SolverState st;
    SolverParameters par;
    std::vector<float2> delta, r, z, p, ap, pre;
    std::vector<float> deltaA, rA, zA, pA, apA, preA;
};

static void bind(Problem& P, int W, int H, float* x, float* A, float* ur, float* cons, float* mask, float wfit, float wreg) {
    const size_t n = (size_t)W * H;
    P.in.N = (unsigned)n; P.in.width = (unsigned)W; P.in.height = (unsigned)H;
    P.in.d_constraints = (float2*)cons;
    std::memset(&P.st, 0, sizeof P.st);
    P.delta.assign(n, make_float2(0, 0)); P.deltaA.assign(n, 0.f);
    P.p.assign(n, make_float2(0, 0)); P.pA.assign(n, 0.f);
    P.pre.assign(n, make_float2(0, 0)); P.preA.assign(n, 0.f);
    P.st.d_delta = P.delta.data(); P.st.d_deltaA = P.deltaA.data();
    P.st.d_x = (float2*)x; P.st.d_A = A; P.st.d_urshape = (float2*)ur; P.st.d_mask = mask;
    P.st.d_p = P.p.data(); P.st.d_pA = P.pA.data();
    P.st.d_precondioner = P.pre.data(); P.st.d_precondionerA = P.preA.data();
    P.par.weightFitting = wfit; P.par.weightRegularizer = wreg;
    P.par.nNonLinearIterations = 1; P.par.nLinIterations = 1;
}

// sum over pixels of evalFDevice
double iw_hand_cost(int W, int H, float* x, float* A, float* ur, float* cons, float* mask, float wfit, float wreg) {
    Problem P; bind(P, W, H, x, A, ur, cons, mask, wfit, wreg);
    double s = 0;
    for (unsigned i = 0; i < P.in.N; ++i) s += (double)evalFDevice(i, P.in, P.st, P.par);
    return s;
}
// evalMinusJTFDevice for every pixel: out_b (2 per pixel), out_bA (1 per pixel)
void iw_hand_minus_jtf(int W, int H, float* x, float* A, float* ur, float* cons, float* mask, float wfit, float wreg,
                       float* out_b, float* out_bA) {
    Problem P; bind(P, W, H, x, A, ur, cons, mask, wfit, wreg);
    for (unsigned i = 0; i < P.in.N; ++i) {
        float bA = 0;
        const float2 b = evalMinusJTFDevice(i, P.in, P.st, P.par, bA);
        out_b[2 * i] = b.x; out_b[2 * i + 1] = b.y; out_bA[i] = bA;
    }
}
// applyJTJDevice for every pixel on the direction (p, pA)
void iw_hand_apply_jtj(int W, int H, float* x, float* A, float* ur, float* cons, float* mask, float wfit, float wreg,
                       const float* p, const float* pA, float* out, float* outA) {
    Problem P; bind(P, W, H, x, A, ur, cons, mask, wfit, wreg);
    std::memcpy(P.p.data(), p, sizeof(float) * 2 * P.in.N);
    std::memcpy(P.pA.data(), pA, sizeof(float) * P.in.N);
    for (unsigned i = 0; i < P.in.N; ++i) {
        float bA = 0;
        const float2 b = applyJTJDevice(i, P.in, P.st, P.par, bA);
        out[2 * i] = b.x; out[2 * i + 1] = b.y; outA[i] = bA;
    }
}
}
