// energy_sfs.cu -- hand-written kernel set for shape_from_shading (BASELINE config "shape_from_shading 640x480 fp32, LM").
//
// All per-pixel math lives in sfs_device.cuh (host/device code, verified on the CPU against the oracle by
// tests/test_sfs_host_parity.py); the kernels here are one-thread-per-pixel grid-stride loops around those functions plus
// the deterministic reductions of common.cuh.  Replaces, for this energy, the generated function map of the reference
// (o.t:2029-2460) and kernels.precompute (solverGPUGaussNewton.t:607-614).
//   precompute   B, dB, flags (pass 1), then the r_gR / r_gC row coefficients (pass 2)       once per linearisation
//   pcgInit      r = -J^T F, pre = guardedInvert(1 | diag), p = pre r, delta = 0, r.p          PCGInit1 :361-397
//   applyJTJ     out = (J^T J [+ CtC]) x, x.out                                                PCGStep1 :421-434
//   cost/modelCost, computeCtC                                                                 :580-592, :666-678, :616-622
// Per PCG step the gather reads p, the 10 row coefficients and the flag byte of a pixel's neighbourhood (L1/L2 hits after
// the first touch): ~53 B/pixel from HBM against the 208+ B/pixel of the cached-Jacobian engine this energy used before.
// Default for this energy (OPTB_GENERIC=1 selects the cached-Jacobian engine instead); parity: tests/test_gpu_generic_energies.py.
#include <type_traits>
#include "solver.h"
#include "sfs_device.cuh"

namespace optb {

__device__ __forceinline__ float sfsGuardedInvert(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }     // CERES variant, solverGPUGaussNewton.t:323-332
__device__ __forceinline__ double sfsGuardedInvert(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }

template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsPrecomputeB(SfsData<T> P) {
    const long N = (long)P.W * P.H;
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock) sfsPrecomputeB(P, (int)(i % P.W), (int)(i / P.W));
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsBuildRows(SfsData<T> P) {
    const long N = (long)P.W * P.H;
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock) sfsBuildRows(P, (int)(i % P.W), (int)(i / P.W));
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsInit(SfsData<T> P, int usePre, ReduceWorkspace ws, int rzSlot, T* __restrict__ delta, T* __restrict__ r,
                                                    T* __restrict__ pre, T* __restrict__ p) {
    const long N = (long)P.W * P.H;
    T acc = T(0);
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock) {
        const int x = (int)(i % P.W), y = (int)(i / P.W);
        T rr = T(0), pr = T(0), pp = T(0);
        if (P.fl[i] & kSfsIncluded) {
            rr = -sfsGather<T, 1>(P, x, y, nullptr);
            T d = usePre ? sfsGather<T, 2>(P, x, y, nullptr) : T(1);     // PCGInit1 order (:379-385): without a preconditioner the constant 1 is inverted too
            pr = sfsGuardedInvert(d);
            pp = pr * rr;
            acc += rr * pp;
        }
        delta[i] = T(0); r[i] = rr; pre[i] = pr; p[i] = pp;
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}
template <class T, bool LM>
__global__ void __launch_bounds__(kBlock) k_sfsApplyJTJ(SfsData<T> P, ReduceWorkspace ws, int dotSlot, const T* __restrict__ x, T* __restrict__ out,
                                                        const T* __restrict__ CtC) {
    const long N = (long)P.W * P.H;
    T acc = T(0);
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock) {
        T o = T(0);
        if (P.fl[i] & kSfsIncluded) {
            o = sfsGather<T, 0>(P, (int)(i % P.W), (int)(i / P.W), x);
            const T xi = x[i];
            if (LM) o += CtC[i] * xi;
            acc += xi * o;
        }
        out[i] = o;
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}
template <class T, bool MODEL>
__global__ void __launch_bounds__(kBlock) k_sfsCost(SfsData<T> P, ReduceWorkspace ws, int slot, const T* __restrict__ delta) {
    const long N = (long)P.W * P.H;
    T acc = T(0);
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock)
        acc += sfsCostAt<T, MODEL>(P, (int)(i % P.W), (int)(i / P.W), delta);
    gridReduceTo<kBlock>((double)acc, ws, slot);
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsCtC(SfsData<T> P, T* __restrict__ CtC, T invRadius) {
    const long N = (long)P.W * P.H;
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock)
        CtC[i] = sfsGather<T, 2>(P, (int)(i % P.W), (int)(i / P.W), nullptr) * invRadius;
}

template <class T>
struct SfsOps : EnergyOps<T> {
    SfsData<T> P{};
    long N = 0; int grid = 1;
    SfsOps(const unsigned* dims) {
        P.W = (int)dims[0]; P.H = (int)dims[1]; N = (long)P.W * P.H;
        this->unknowns.resize(1);
        this->unknowns[0].elements = N; this->unknowns[0].channels = 1;
        this->layoutUnknowns();
        this->usePreconditioner = false;   // the spec has no UsePreconditioner(): default false (o.t:214)
        OPT_CUDA_CHECK(cudaMalloc((void**)&P.B, sizeof(T) * N)); OPT_CUDA_CHECK(cudaMalloc((void**)&P.dB, sizeof(T) * 3 * N));
        OPT_CUDA_CHECK(cudaMalloc((void**)&P.cR, sizeof(T) * 5 * N)); OPT_CUDA_CHECK(cudaMalloc((void**)&P.cC, sizeof(T) * 5 * N));
        OPT_CUDA_CHECK(cudaMalloc((void**)&P.fl, N)); OPT_CUDA_CHECK(cudaMemset(P.fl, 0, N));
        int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        long g = (N + kBlock - 1) / kBlock; if (g > (long)sms * 8) g = (long)sms * 8; if (g > kMaxPartials) g = kMaxPartials; if (g < 1) g = 1;
        grid = (int)g;
    }
    ~SfsOps() override { cudaFree(P.B); cudaFree(P.dB); cudaFree(P.cR); cudaFree(P.cC); cudaFree(P.fl); }
    const char* name() const override { return "shape_from_shading"; }
    unsigned long long captureKey() const override { return EnergyOps<T>::hashValue(P); }
    void bind(void** q) override {   // positions: shape_from_shading.t:5-28 (all scalar Params are float, cast to opt_float on use, o.t:1699)
        auto f = [&](int k) { return (T) * (const float*)q[k]; };
        P.wp = (T)sqrt((double)f(0)); P.ws = (T)sqrt((double)f(1)); P.wg = (T)sqrt((double)f(2));
        if (std::is_same<T, float>::value) { P.wp = (T)sqrtf(*(const float*)q[0]); P.ws = (T)sqrtf(*(const float*)q[1]); P.wg = (T)sqrtf(*(const float*)q[2]); }
        P.fx = f(3); P.fy = f(4); P.ux = f(5); P.uy = f(6);
        for (int i = 1; i <= 9; ++i) P.L[i] = f(6 + i);
        this->unknowns[0].user = q[16];
        P.X = (const T*)q[16]; P.D = (const T*)q[17]; P.Im = (const T*)q[18]; P.eR = (const unsigned char*)q[19]; P.eC = (const unsigned char*)q[20];
    }
    void precompute(const LaunchCtx& c) override {
        k_sfsPrecomputeB<T><<<grid, kBlock, 0, c.stream>>>(P);
        k_sfsBuildRows<T><<<grid, kBlock, 0, c.stream>>>(P);
    }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        k_sfsInit<T><<<grid, kBlock, 0, c.stream>>>(P, this->usePreconditioner ? 1 : 0, c.ws, rzSlot, delta, r, pre, p);
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan*) override {
        if (CtC) k_sfsApplyJTJ<T, true><<<grid, kBlock, 0, c.stream>>>(P, c.ws, dotSlot, x, out, CtC);
        else k_sfsApplyJTJ<T, false><<<grid, kBlock, 0, c.stream>>>(P, c.ws, dotSlot, x, out, nullptr);
    }
    void cost(const LaunchCtx& c) override { k_sfsCost<T, false><<<grid, kBlock, 0, c.stream>>>(P, c.ws, kCost, nullptr); }
    void modelCost(const LaunchCtx& c, const T* delta) override { k_sfsCost<T, true><<<grid, kBlock, 0, c.stream>>>(P, c.ws, kModelCost, delta); }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override { k_sfsCtC<T><<<grid, kBlock, 0, c.stream>>>(P, CtC, invRadius); }
};

template <class T> EnergyOps<T>* makeSfs(const unsigned* dims) { return new SfsOps<T>(dims); }
template EnergyOps<float>* makeSfs<float>(const unsigned*);
template EnergyOps<double>* makeSfs<double>(const unsigned*);

}  // namespace optb
