// ref_arap_shim.cpp -- host-side harness around the REFERENCE's own hand-written arap_mesh_deformation equations
// (/root/reference/examples/arap_mesh_deformation/src/WarpingSolverEquations.h: evalFDevice :14, evalMinusJTFDevice :52,
// applyJTJDevice :131), compiled for the CPU straight from the reference tree (nothing is copied into this repository).
// TEST INFRASTRUCTURE: pins the CPU oracle's 3-D ARAP energy the way ref_iw_shim.cpp pins image_warping.
// The device qualifiers reduce to ignored attributes under g++; the device intrinsics the headers mention get host stubs.
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <vector>
static inline float __shfl_down(float v, int, int) { return v; }
static inline void __syncthreads() {}
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, sizeof f); return f; }   // SolverUtil.h: MINF = __int_as_float(0xff800000)
#include "WarpingSolverState.h"
#include "WarpingSolverParameters.h"
#include "WarpingSolverEquations.h"

extern "C" {
// The hand solver walks per-vertex neighbour lists (both directions of every mesh edge are listed, as the example builds them);
// cost_ref = sum_v evalFDevice = 2 x Opt's cost with weightFitting = w_fitSqrt^2 etc.; b = -J^T F in that 2x convention.
void ref_arap_eval(int N, float3* x, float3* a, float3* urshape, float3* target, int* numNeighbours, int* neighbourIdx, int* neighbourOffset,
                   float weightFit, float weightReg, double* cost, float3* b, float3* bA) {
    SolverInput in; in.N = N; in.d_numNeighbours = numNeighbours; in.d_neighbourIdx = neighbourIdx; in.d_neighbourOffset = neighbourOffset;
    SolverParameters pr; pr.weightFitting = weightFit; pr.weightRegularizer = weightReg; pr.nNonLinearIterations = 0; pr.nLinIterations = 0;
    std::vector<float3> delta(N), deltaA(N), pre(N), preA(N);
    SolverState st{}; st.d_x = x; st.d_a = a; st.d_urshape = urshape; st.d_target = target; st.d_delta = delta.data(); st.d_deltaA = deltaA.data();
    st.d_precondioner = pre.data(); st.d_precondionerA = preA.data();
    double c = 0;
    for (int i = 0; i < N; ++i) { c += evalFDevice(i, in, st, pr); float3 ba; b[i] = evalMinusJTFDevice(i, in, st, pr, ba); bA[i] = ba; }
    *cost = c;
}
void ref_arap_applyjtj(int N, float3* x, float3* a, float3* urshape, float3* target, int* numNeighbours, int* neighbourIdx, int* neighbourOffset,
                       float weightFit, float weightReg, float3* p, float3* pA, float3* out, float3* outA) {
    SolverInput in; in.N = N; in.d_numNeighbours = numNeighbours; in.d_neighbourIdx = neighbourIdx; in.d_neighbourOffset = neighbourOffset;
    SolverParameters pr; pr.weightFitting = weightFit; pr.weightRegularizer = weightReg; pr.nNonLinearIterations = 0; pr.nLinIterations = 0;
    SolverState st{}; st.d_x = x; st.d_a = a; st.d_urshape = urshape; st.d_target = target; st.d_p = p; st.d_pA = pA;
    for (int i = 0; i < N; ++i) { float3 oa; out[i] = applyJTJDevice(i, in, st, pr, oa); outA[i] = oa; }
}
}
