// ref_iw_shim.cpp -- host-side harness around the REFERENCE's own hand-written image_warping equations
// (/root/reference/examples/image_warping/src/WarpingSolverEquations.h: evalFDevice :9,
// evalMinusJTFDevice :45, applyJTJDevice :207), compiled for the CPU straight from the reference tree
// (nothing is copied into this repository).  TEST INFRASTRUCTURE: pins the CPU oracle.
// The device qualifiers reduce to ignored attributes under g++; the two device intrinsics the included
// utility header mentions (never reached from the three functions used here) get host stubs.
#include <cuda_runtime.h>
#include <cmath>
#include <vector>
static inline float __shfl_down(float v, int, int) { return v; }
static inline void __syncthreads() {}
#include "WarpingSolverState.h"
#include "WarpingSolverParameters.h"
#include "WarpingSolverEquations.h"

extern "C" {
// cost_ref = sum_i evalFDevice (NOTE: the hand solver's convention is sum w*r^2 with unsquared weights, = 2x Opt's cost)
// b = -J^T F in the hand solver's 2x convention; pre/preA = its inverted Jacobi preconditioner
void ref_iw_eval(int W, int H, float2* x, float* A, float2* urshape, float* mask, float2* constraints, float weightFit, float weightReg,
                 double* cost, float2* b, float* bA, float2* pre, float* preA) {
    SolverInput in; in.N = W * H; in.width = W; in.height = H; in.d_constraints = constraints;
    SolverParameters pr; pr.weightFitting = weightFit; pr.weightRegularizer = weightReg; pr.nNonLinearIterations = 0; pr.nLinIterations = 0;
    std::vector<float2> delta(W * H); std::vector<float> deltaA(W * H);
    SolverState st{}; st.d_x = x; st.d_A = A; st.d_urshape = urshape; st.d_mask = mask; st.d_delta = delta.data(); st.d_deltaA = deltaA.data();
    st.d_precondioner = pre; st.d_precondionerA = preA;
    double c = 0;
    for (int i = 0; i < W * H; ++i) {
        if (mask[i] == 0) { c += evalFDevice(i, in, st, pr); float a; b[i] = evalMinusJTFDevice(i, in, st, pr, a); bA[i] = a; }
        else { b[i] = make_float2(0, 0); bA[i] = 0; }
    }
    *cost = c;
}
void ref_iw_applyjtj(int W, int H, float2* x, float* A, float2* urshape, float* mask, float2* constraints, float weightFit, float weightReg,
                     float2* p, float* pA, float2* out, float* outA) {
    SolverInput in; in.N = W * H; in.width = W; in.height = H; in.d_constraints = constraints;
    SolverParameters pr; pr.weightFitting = weightFit; pr.weightRegularizer = weightReg; pr.nNonLinearIterations = 0; pr.nLinIterations = 0;
    SolverState st{}; st.d_x = x; st.d_A = A; st.d_urshape = urshape; st.d_mask = mask; st.d_p = p; st.d_pA = pA;
    for (int i = 0; i < W * H; ++i) {
        if (mask[i] == 0) { float a; out[i] = applyJTJDevice(i, in, st, pr, a); outA[i] = a; }
        else { out[i] = make_float2(0, 0); outA[i] = 0; }
    }
}
}
