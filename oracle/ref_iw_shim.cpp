// ref_iw_shim.cpp -- host-side harness around the REFERENCE's own hand-written image_warping equations
// (/root/reference/examples/image_warping/src/WarpingSolverEquations.h: evalFDevice :9,
// evalMinusJTFDevice :45, applyJTJDevice :207), compiled for the CPU straight from the reference tree
// (nothing is copied into this repository).  TEST INFRASTRUCTURE: pins the CPU oracle.
// The device qualifiers reduce to ignored attributes under g++; the two device intrinsics the included
// utility header mentions (never reached from the three functions used here) get host stubs.
#include <cuda_runtime.h>
#include <cmath>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
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
// Matrix-free PCG iterations timed on the host cores, built from the reference's OWN per-pixel functions
// (evalMinusJTFDevice for the right-hand side and its inverted Jacobi preconditioner, applyJTJDevice for A p): the same call
// sequence as the hand-written solver's PCGInit_Kernel1/2 and PCGStep_Kernel1/2/3 (WarpingSolver.cu:87-260), with the
// kernels' thread grids replaced by OpenMP loops over the pixels.  This is bench.py's CPU baseline of kind "reference":
// the reference's own arithmetic for PCGStep1, run on the CPU.  Returns the seconds spent in `steps` iterations
// (the initialisation is not timed); *threads = OpenMP threads used; *rz = r.z after the last step (so the loop cannot be
// optimised away and the caller can see that it converges).
double ref_iw_pcg(int W, int H, float2* x, float* A, float2* urshape, float* mask, float2* constraints, float weightFit, float weightReg,
                  int steps, int* threads, double* rzOut) {
    // *threads > 0 on entry: use exactly that many OpenMP threads (the caller passes the cores its cgroup / affinity mask really
    // grants: omp_get_max_threads() reports every core of the host, 128 on a box whose container may run on 16)
#ifdef _OPENMP
    if (threads && *threads > 0) omp_set_num_threads(*threads);
#endif
    const int N = W * H;
    SolverInput in; in.N = N; in.width = W; in.height = H; in.d_constraints = constraints;
    SolverParameters pr; pr.weightFitting = weightFit; pr.weightRegularizer = weightReg; pr.nNonLinearIterations = 0; pr.nLinIterations = 0;
    // work vectors: allocated untouched and first written inside the parallel loops below (first-touch page placement follows the
    // static schedule every later loop uses)
    struct Buf2 { float2* v; explicit Buf2(size_t n) : v((float2*)malloc(n * sizeof(float2))) {} ~Buf2() { free(v); } float2* data() { return v; } float2& operator[](size_t i) { return v[i]; } };
    struct Buf1 { float* v; explicit Buf1(size_t n) : v((float*)malloc(n * sizeof(float))) {} ~Buf1() { free(v); } float* data() { return v; } float& operator[](size_t i) { return v[i]; } };
    Buf2 delta(N), r(N), z(N), p(N), Ap(N), pre(N);
    Buf1 deltaA(N), rA(N), zA(N), pA(N), ApA(N), preA(N);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        delta[i] = r[i] = z[i] = p[i] = Ap[i] = pre[i] = make_float2(0, 0);
        deltaA[i] = rA[i] = zA[i] = pA[i] = ApA[i] = preA[i] = 0.f;
    }
    SolverState st{}; st.d_x = x; st.d_A = A; st.d_urshape = urshape; st.d_mask = mask; st.d_delta = delta.data(); st.d_deltaA = deltaA.data();
    st.d_precondioner = pre.data(); st.d_precondionerA = preA.data(); st.d_p = p.data(); st.d_pA = pA.data();
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (threads) *threads = nt;
    double rz = 0;
#pragma omp parallel for reduction(+ : rz) schedule(static)
    for (int i = 0; i < N; ++i) {
        float2 b = make_float2(0, 0); float bA = 0;
        if (mask[i] == 0) { SolverInput in_ = in; SolverState st_ = st; SolverParameters pr_ = pr; b = evalMinusJTFDevice(i, in_, st_, pr_, bA); }
        else { pre[i] = make_float2(0, 0); preA[i] = 0; delta[i] = make_float2(0, 0); deltaA[i] = 0; }
        r[i] = b; rA[i] = bA;
        p[i] = make_float2(pre[i].x * b.x, pre[i].y * b.y); pA[i] = preA[i] * bA;
        rz += (double)b.x * p[i].x + (double)b.y * p[i].y + (double)bA * pA[i];
    }
    double t0 = 0, t1 = 0;
#ifdef _OPENMP
    t0 = omp_get_wtime();
#else
    t0 = (double)clock() / CLOCKS_PER_SEC;
#endif
    for (int it = 0; it < steps; ++it) {
        double pAp = 0;
#pragma omp parallel for reduction(+ : pAp) schedule(static)
        for (int i = 0; i < N; ++i) {                                   // PCGStep_Kernel1: Ap = J^T J p, p.Ap
            float2 o = make_float2(0, 0); float oA = 0;
            if (mask[i] == 0) { SolverInput in_ = in; SolverState st_ = st; SolverParameters pr_ = pr; o = applyJTJDevice(i, in_, st_, pr_, oA); }
            Ap[i] = o; ApA[i] = oA;
            pAp += (double)p[i].x * o.x + (double)p[i].y * o.y + (double)pA[i] * oA;
        }
        const float alpha = pAp > 0 ? (float)(rz / pAp) : 0.f;
        double rzNew = 0;
#pragma omp parallel for reduction(+ : rzNew) schedule(static)
        for (int i = 0; i < N; ++i) {                                   // PCGStep_Kernel2: delta, r, z, r.z
            delta[i].x += alpha * p[i].x; delta[i].y += alpha * p[i].y; deltaA[i] += alpha * pA[i];
            r[i].x -= alpha * Ap[i].x; r[i].y -= alpha * Ap[i].y; rA[i] -= alpha * ApA[i];
            z[i] = make_float2(pre[i].x * r[i].x, pre[i].y * r[i].y); zA[i] = preA[i] * rA[i];
            rzNew += (double)z[i].x * r[i].x + (double)z[i].y * r[i].y + (double)zA[i] * rA[i];
        }
        const float beta = rz > 0 ? (float)(rzNew / rz) : 0.f;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i) {                                   // PCGStep_Kernel3: p = z + beta p
            p[i].x = z[i].x + beta * p[i].x; p[i].y = z[i].y + beta * p[i].y; pA[i] = zA[i] + beta * pA[i];
        }
        rz = rzNew;
    }
#ifdef _OPENMP
    t1 = omp_get_wtime();
#else
    t1 = (double)clock() / CLOCKS_PER_SEC;
#endif
    if (rzOut) *rzOut = rz;
    return t1 - t0;
}
}
