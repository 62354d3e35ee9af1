// common.cuh -- shared device/host helpers for libOpt.so (sm_100a).
//
// Replaces the reference's GPU runtime utilities (API/src/util.t): error checking
// (`cd`/`checkedLaunch`, util.t:739-768), launch geometry (util.t:782-843) and the reduction
// primitive (`warpReduce` + one global atomicAdd per warp, util.t:526-623 /
// solverGPUGaussNewton.t:312-317).  The reduction here is different by design: every CTA writes ONE
// double partial, the last CTA to finish sums the partials in index order, so results are
// deterministic run-to-run and no same-address float atomics are issued.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define OPT_CUDA_CHECK(call)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            fprintf(stderr, "Cuda reported error %d: %s\nIn call: %s\nIn file: %s:%d\n", (int)_e,  \
                    cudaGetErrorString(_e), #call, __FILE__, __LINE__);                            \
            exit((int)_e); /* same contract as the reference: print and exit(code), util.t:739-752 */ \
        }                                                                                          \
    } while (0)

namespace optb {

constexpr int kBlock = 256;          // threads per CTA for all solver kernels
constexpr int kMaxPartials = 4096;   // upper bound on CTAs that contribute to one reduction

// Device-resident scalars of one plan (the reference keeps 6 separate 1-element allocations,
// solverGPUGaussNewton.t:1274-1280).  Kept in double; kernels cast to opt_float on use.
enum Scalar : int {
    kRz0 = 0,       // r.z ping  (scanAlphaNumerator / scanBetaNumerator alternate by PCG-iteration parity)
    kRz1 = 1,       // r.z pong
    kPAp = 2,       // p.Ap      (scanAlphaDenominator)
    kCost = 3,      // scratch   (cost)
    kModelCost = 4, // modelCost
    kQ = 5,         // q (LM)
    kQPrev = 6,     // previous q (LM zeta test on device)
    kFlags = 7,     // device-side control word (LM early-out), stored as a double 0/1
    kNumScalars = 16
};

struct ReduceWorkspace {
    double* partials;        // [kNumScalars][kMaxPartials]
    unsigned int* counters;  // [kNumScalars]  arrival counters, self-resetting
    double* scalars;         // [kNumScalars]
};

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of `v` (every thread calls), then grid-wide deterministic finish:
// partial -> partials[slot][blockIdx]; last CTA sums gridDim.x partials in order -> scalars[slot].
// Must be called by all threads of all CTAs of the launch exactly once per slot.
template <int NT>
__device__ __forceinline__ void gridReduceTo(double v, const ReduceWorkspace& ws, int slot) {
    __shared__ double s_warp[NT / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warpSum(v);
    if (lane == 0) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < NT / 32 ? s_warp[lane] : 0.0;
        t = warpSum(t);
        if (lane == 0) {
            double* part = ws.partials + (size_t)slot * kMaxPartials;
            part[blockIdx.x] = t;
            __threadfence();
            unsigned int prev = atomicAdd(ws.counters + slot, 1u);
            s_last = (prev == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        const volatile double* part = ws.partials + (size_t)slot * kMaxPartials;
        double t = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += NT) t += part[i];
        // fixed summation tree: thread-strided partial sums, then the same shuffle tree every run
        t = warpSum(t);
        if (lane == 0) s_warp[wid] = t;
        __syncthreads();
        if (wid == 0) {
            double u = lane < NT / 32 ? s_warp[lane] : 0.0;
            u = warpSum(u);
            if (lane == 0) { ws.scalars[slot] = u; ws.counters[slot] = 0u; }
        }
    }
    __syncthreads();   // s_warp / s_last reusable by a second call in the same kernel
}

inline int divUp(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace optb
