// ORACLE (test infrastructure): host-side driver around the reference's HAND-DERIVED shading term of shape_from_shading
// (examples/shape_from_shading/src/SFSSolverUtil.h: calShading2depthGradHelper :59-195 -- the spherical-harmonics shading
// error B(x, y) - I(x, y) of a pixel and its derivatives with respect to the three depth samples it reads), compiled from
// the reference tree where it lies (`make -C oracle hand` -> oracle/_ref/libsfs_hand.so, nothing copied).  The test
// compares it with the value and the gradient image of the ComputedArray `B_I_comp` that the oracle derives from
// shape_from_shading.t by dual-number AD (tests/test_oracle_hand_equations.py).
#include "hand_shims.h"

#include REF_UTIL_HEADER

extern "C" {
// out[4 * (y * W + x) + {0, 1, 2, 3}] = d/d X(x-1, y), d/d X(x, y), d/d X(x, y-1), B - I   for 1 <= x < W, 1 <= y < H
void sfs_hand_shading(int W, int H, const float* X, float* intensity, float* light9, float fx, float fy, float ux, float uy, float* out) {
    SolverInput in;
    std::memset(&in, 0, sizeof in);
    in.N = (unsigned)(W * H); in.width = (unsigned)W; in.height = (unsigned)H;
    in.d_targetIntensity = intensity; in.d_litcoeff = light9;
    in.calibparams.fx = fx; in.calibparams.fy = fy; in.calibparams.ux = ux; in.calibparams.uy = uy;
    for (int y = 1; y < H; ++y)
        for (int x = 1; x < W; ++x) {
            const float d0 = X[y * W + x - 1], d1 = X[y * W + x], d2 = X[(y - 1) * W + x];
            const float4 r = calShading2depthGradHelper(d0, d1, d2, x, y, in);
            float* o = out + 4 * (y * W + x);
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
        }
}
}
