// ORACLE (test infrastructure): host-side driver around the reference's HAND-DERIVED volumetric_mesh_deformation equations
// (examples/volumetric_mesh_deformation/src/WarpingSolverEquations.h: evalFDevice :35-70, evalMinusJTFDevice :76-146,
// applyJTJDevice :152-..., rotations in RotationHelper.h; a (dx+1) x (dy+1) x (dz+1) node lattice, z fastest), compiled from the reference tree where it lies
// (`make -C oracle hand` -> oracle/_ref/libvol_hand.so, nothing copied).  Same purpose and scaling as iw_hand_host.cpp:
// the hand solver minimises sum w e^2, Thallo 1/2 sum (sqrt(w) e)^2.
#include "hand_shims.h"

#include REF_STATE_HEADER
#include REF_EQUATIONS_HEADER

extern "C" {

struct Problem {
    SolverInput in;
    SolverState st;
    SolverParameters par;
    std::vector<float3> delta, deltaA, p, pA, pre, preA;
};

// nodes[3] = lattice nodes along the reference's (x, y, z), x slowest: the reference stores cells, i.e. nodes - 1
static void bind(Problem& P, int N, float* x, float* a, float* target, float* ur, const int* nodes, float wfit, float wreg) {
    std::memset(&P.in, 0, sizeof P.in); std::memset(&P.st, 0, sizeof P.st);
    P.in.N = (unsigned)N; P.in.dims = make_int3(nodes[0] - 1, nodes[1] - 1, nodes[2] - 1);
    const float3 z = make_float3(0, 0, 0);
    P.delta.assign(N, z); P.deltaA.assign(N, z); P.p.assign(N, z); P.pA.assign(N, z); P.pre.assign(N, z); P.preA.assign(N, z);
    P.st.d_delta = P.delta.data(); P.st.d_deltaA = P.deltaA.data();
    P.st.d_x = (float3*)x; P.st.d_a = (float3*)a; P.st.d_target = (float3*)target; P.st.d_urshape = (float3*)ur;
    P.st.d_p = P.p.data(); P.st.d_pA = P.pA.data();
    P.st.d_precondioner = P.pre.data(); P.st.d_precondionerA = P.preA.data();
    P.par.weightFitting = wfit; P.par.weightRegularizer = wreg; P.par.nNonLinearIterations = 1; P.par.nLinIterations = 1;
}

double vol_hand_cost(int N, float* x, float* a, float* target, float* ur, const int* nodes, float wfit, float wreg) {
    Problem P; bind(P, N, x, a, target, ur, nodes, wfit, wreg);
    double s = 0;
    for (int i = 0; i < N; ++i) s += (double)evalFDevice((unsigned)i, P.in, P.st, P.par);
    return s;
}
void vol_hand_minus_jtf(int N, float* x, float* a, float* target, float* ur, const int* nodes, float wfit, float wreg,
                         float* out_b, float* out_bA) {
    Problem P; bind(P, N, x, a, target, ur, nodes, wfit, wreg);
    for (int i = 0; i < N; ++i) {
        float3 bA;
        const float3 b = evalMinusJTFDevice((unsigned)i, P.in, P.st, P.par, bA);
        out_b[3 * i] = b.x; out_b[3 * i + 1] = b.y; out_b[3 * i + 2] = b.z;
        out_bA[3 * i] = bA.x; out_bA[3 * i + 1] = bA.y; out_bA[3 * i + 2] = bA.z;
    }
}
void vol_hand_apply_jtj(int N, float* x, float* a, float* target, float* ur, const int* nodes, float wfit, float wreg,
                         const float* p, const float* pA, float* out, float* outA) {
    Problem P; bind(P, N, x, a, target, ur, nodes, wfit, wreg);
    std::memcpy(P.p.data(), p, sizeof(float) * 3 * N);
    std::memcpy(P.pA.data(), pA, sizeof(float) * 3 * N);
    for (int i = 0; i < N; ++i) {
        float3 bA;
        const float3 b = applyJTJDevice((unsigned)i, P.in, P.st, P.par, bA);
        out[3 * i] = b.x; out[3 * i + 1] = b.y; out[3 * i + 2] = b.z;
        outA[3 * i] = bA.x; outA[3 * i + 1] = bA.y; outA[3 * i + 2] = bA.z;
    }
}
}
