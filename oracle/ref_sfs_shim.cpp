// ref_sfs_shim.cpp -- host-side harness around the REFERENCE's own hand-written shape_from_shading equations
// (/root/reference/examples/shape_from_shading/src/SFSSolverEquations.h: evalFDevice :64, evalMinusJTFDevice :237,
// applyJTJDevice :473; helpers SFSSolverUtil.h), compiled for the CPU straight from the reference tree (nothing is copied into this
// repository).  TEST INFRASTRUCTURE: pins the CPU oracle's shape_from_shading energy the way ref_iw_shim.cpp pins
// image_warping.  Call sequence = the reference's own dump stub (SFSSolver.cu:409-463, solveSFSEvalCurrentCostJTFPreAndJTJStub):
// Precompute, per-pixel cost / -J^T F / preconditioner, p = the caller's vector, J^T J p.  Precompute_Kernel (SFSSolver.cu:317-346)
// lives in the .cu (kernel launch syntax, not host-compilable), so its dozen lines are restated in precompute() below from the same
// header functions it calls (calShading2depthGradCompute, IsValidPoint).
// Convention map (as in the other shims): the hand solver minimises sum w r^2 with UNSQUARED weights, i.e. cost_hand = 2 cost_opt,
// evalMinusJTFDevice = -(1/2) d(cost_hand)/dx = -J^T F (Opt convention, weights squared-in), applyJTJDevice = J^T J p.
// Domain caveat the tests account for: the hand solver only evaluates pixels with 1 < x < W-5, 1 < y < H-5 (Equations.h:83,:258),
// Opt's spec uses InBoundsExpanded(0,0,1) -- the two agree away from the image border.
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>
static inline float __shfl_down(float v, int, int) { return v; }
static inline void __syncthreads() {}
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, sizeof f); return f; }   // SolverUtil.h: MINF = __int_as_float(0xff800000)
#include "SFSSolverState.h"
#include "SFSSolverParameters.h"
#include "SFSSolverEquations.h"

namespace {
struct Problem {
    SolverInput in{}; SolverState st{}; SolverParameters pr{};
    std::vector<float> B, B0, B1, B2; std::vector<char> guard;
    Problem(int W, int H, float* X, float* D, float* Im, unsigned char* eR, unsigned char* eC, const float* prm /* w_p w_s w_g fx fy ux uy L1..L9 */) {
        in.N = (unsigned)W * H; in.width = W; in.height = H;
        in.d_targetIntensity = Im; in.d_targetDepth = D; in.d_maskEdgeMapR = eR; in.d_maskEdgeMapC = eC;
        in.d_litcoeff = const_cast<float*>(prm + 7);
        in.calibparams.fx = prm[3]; in.calibparams.fy = prm[4]; in.calibparams.ux = prm[5]; in.calibparams.uy = prm[6];
        pr.weightFitting = prm[0]; pr.weightRegularizer = prm[1]; pr.weightShading = prm[2]; pr.weightShadingStart = prm[2];
        pr.weightPrior = 0; pr.weightShadingIncrement = 0; pr.weightBoundary = 0; pr.nNonLinearIterations = 0; pr.nLinIterations = 0; pr.nPatchIterations = 0;
        st.d_x = X;
        B.assign(in.N, 0.f); B0 = B; B1 = B; B2 = B; guard.assign(in.N, 0);
        st.B_I = B.data(); st.B_I_dx0 = B0.data(); st.B_I_dx1 = B1.data(); st.B_I_dx2 = B2.data(); st.pguard = reinterpret_cast<bool*>(guard.data());
        precompute();
    }
    void precompute() {   // SFSSolver.cu:317-346
        const int W = in.width, H = in.height;
        for (int posy = 0; posy < H; ++posy) for (int posx = 0; posx < W; ++posx) {
            const int x = posy * W + posx;
            if (posx > 0 && posy > 0 && posx < W - 2 && posy < H - 2) {
                float4 t = calShading2depthGradCompute(st, posx, posy, in);
                st.B_I_dx0[x] = t.x; st.B_I_dx1[x] = t.y; st.B_I_dx2[x] = t.z; st.B_I[x] = t.w;
                const float d = readX(st, posy, posx, W), d0 = readX(st, posy, posx - 1, W), d1 = readX(st, posy, posx + 1, W), d2 = readX(st, posy - 1, posx, W), d3 = readX(st, posy + 1, posx, W);
                st.pguard[x] = IsValidPoint(d) && IsValidPoint(d0) && IsValidPoint(d1) && IsValidPoint(d2) && IsValidPoint(d3)
                    && std::fabs(d - d0) < DEPTH_DISCONTINUITY_THRE && std::fabs(d - d1) < DEPTH_DISCONTINUITY_THRE
                    && std::fabs(d - d2) < DEPTH_DISCONTINUITY_THRE && std::fabs(d - d3) < DEPTH_DISCONTINUITY_THRE;
            }
        }
    }
};
}  // namespace

extern "C" {
// per-pixel cost (hand convention: sum w r^2), -J^T F and the hand solver's preconditioner value
void ref_sfs_eval(int W, int H, float* X, float* D, float* Im, unsigned char* eR, unsigned char* eC, const float* prm, float* cost, float* minusJtF, float* pre) {
    Problem P(W, H, X, D, Im, eR, eC, prm);
    for (int i = 0; i < W * H; ++i) {
        float p = 1.0f;
        cost[i] = evalFDevice(i, P.in, P.st, P.pr);
        minusJtF[i] = evalMinusJTFDevice((unsigned)i, P.in, P.st, P.pr, p);
        pre[i] = p;
    }
}
void ref_sfs_applyjtj(int W, int H, float* X, float* D, float* Im, unsigned char* eR, unsigned char* eC, const float* prm, float* p, float* out) {
    Problem P(W, H, X, D, Im, eR, eC, prm);
    P.st.d_p = p;
    for (int i = 0; i < W * H; ++i) out[i] = applyJTJDevice((unsigned)i, P.in, P.st, P.pr);
}
}
