// ref_poisson_shim.cpp -- host-side harness around the REFERENCE's own hand-written poisson_image_editing equations
// (/root/reference/examples/poisson_image_editing/src/WarpingSolverEquations.h: evalFDevice :12, evalMinusJTFDevice :47,
// applyJTJDevice :85), compiled for the CPU straight from the reference tree (nothing is copied into this repository).
// TEST INFRASTRUCTURE: pins the CPU oracle's poisson_image_editing energy (BASELINE.json configs[0]) the way
// ref_iw_shim.cpp pins image_warping.  The hand solver's kernels (WarpingSolver.cu) call the three functions only for pixels with
// mask == 0; the loops below do the same.
// Convention map: the hand solver counts every neighbour pair ONCE per pixel, Opt's gather counts the residual centred at the pixel
// and the one centred at the neighbour (o.t:1996-2006), so  sum evalFDevice = 2 cost_opt,  evalMinusJTFDevice = -(1/2) J^T F,
// applyJTJDevice = (1/2) J^T J p  (the same Gauss-Newton step; the hand solver's preconditioner is the identity, as UsePreconditioner(false)).
#include <cuda_runtime.h>
#include <cmath>
#include <limits>
#include <vector>
static inline float __shfl_down(float v, int, int) { return v; }
static inline void __syncthreads() {}
#include "WarpingSolverState.h"
#include "WarpingSolverParameters.h"
#include "WarpingSolverEquations.h"

extern "C" {
void ref_poisson_eval(int W, int H, float4* x, float4* target, float* mask, double* cost, float4* minusJtF, float4* pre) {
    SolverInput in; in.N = W * H; in.width = W; in.height = H;
    SolverParameters pr{};
    std::vector<float4> delta(W * H);
    SolverState st{}; st.d_x = x; st.d_target = target; st.d_mask = mask; st.d_delta = delta.data(); st.d_precondioner = pre;
    double c = 0;
    for (int i = 0; i < W * H; ++i) {
        c += evalFDevice(i, in, st, pr);                                    // tests the mask itself (:16)
        if (mask[i] == 0) minusJtF[i] = evalMinusJTFDevice(i, in, st, pr);
        else { minusJtF[i] = make_float4(0, 0, 0, 0); pre[i] = make_float4(1, 1, 1, 1); }
    }
    *cost = c;
}
void ref_poisson_applyjtj(int W, int H, float* mask, float4* p, float4* out) {
    SolverInput in; in.N = W * H; in.width = W; in.height = H;
    SolverParameters pr{};
    SolverState st{}; st.d_mask = mask; st.d_p = p;
    for (int i = 0; i < W * H; ++i) out[i] = mask[i] == 0 ? applyJTJDevice(i, in, st, pr) : make_float4(0, 0, 0, 0);
}
}
