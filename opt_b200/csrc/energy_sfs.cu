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
#include <algorithm>
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
// PCGStep1, tiled: the same operator as k_sfsApplyJTJ in two passes over a 32x8 tile held in shared memory.  The one-pass gather
// re-forms every row's  y = row . p  once per unknown that row touches (5x for the shading rows, 5x3 for the smoothness rows): ~260
// L1/L2 loads per pixel per PCG step, which is what bounds it (46 us per step at 640x480, measured).  Here p (tile + 2-pixel ring),
// then the five row values and the ten shading-row coefficients of every row centre in tile + 1-pixel ring are staged once; the gather
// for unknown u reads them from shared memory.  HBM/L2 traffic per pixel: p, 10 coefficients, the flag byte, out (+CtC).
// Same products in the same order as sfsGather<T, 0> (rows with a zero coefficient contribute an exact 0 instead of being skipped).
constexpr int kSfsTW = 32, kSfsTH = 8;
constexpr int kSfsRW = kSfsTW + 2, kSfsRH = kSfsTH + 2, kSfsPW = kSfsTW + 4, kSfsPH = kSfsTH + 4;
static_assert(kSfsTW * kSfsTH == kBlock, "one thread per tile pixel");
template <class T, bool LM>
__global__ void __launch_bounds__(kBlock) k_sfsApplyJTJTiled(SfsData<T> P, ReduceWorkspace ws, int dotSlot, const T* __restrict__ x, T* __restrict__ out,
                                                             const T* __restrict__ CtC, int tilesX, int numTiles) {
    __shared__ T sp[kSfsPH][kSfsPW];
    __shared__ T sy[5][kSfsRH][kSfsRW];        // yR, yC, yS.x, yS.y, yS.z of the row centred at the cell
    __shared__ T sc[10][kSfsRH][kSfsRW];       // cR[0..4], cC[0..4] of the row centred at the cell
    __shared__ T sAx[kSfsPW], sAy[kSfsPH];     // (x - u_x) / f_x, (y - u_y) / f_y for the columns / rows of the staged region
    const long N = (long)P.W * P.H;
    const int tid = threadIdx.x, tx = tid % kSfsTW, ty = tid / kSfsTW;
    T acc = T(0);
    for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const int x0 = (tile % tilesX) * kSfsTW, y0 = (tile / tilesX) * kSfsTH;
        __syncthreads();   // the previous tile's gather is done with the staging buffers
        for (int c = tid; c < kSfsPH * kSfsPW; c += kBlock) {
            const int ly = c / kSfsPW, lx = c - ly * kSfsPW, gx = x0 - 2 + lx, gy = y0 - 2 + ly;
            sp[ly][lx] = (gx >= 0 && gx < P.W && gy >= 0 && gy < P.H) ? x[(long)gy * P.W + gx] : T(0);
        }
        // back-projection factors of the tile's columns / rows (+2 ring): 36 + 12 divisions per tile instead of 6 per row cell and 2 per pixel
        if (tid < kSfsPW) sAx[tid] = sfsAx(P, x0 - 2 + tid);
        else if (tid >= 64 && tid < 64 + kSfsPH) sAy[tid - 64] = sfsAy(P, y0 - 2 + (tid - 64));
        __syncthreads();
        for (int c = tid; c < kSfsRH * kSfsRW; c += kBlock) {
            const int ly = c / kSfsRW, lx = c - ly * kSfsRW, gx = x0 - 1 + lx, gy = y0 - 1 + ly;
            T y[5] = { T(0), T(0), T(0), T(0), T(0) }, co[10] = { T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0) };
            if (gx >= 0 && gx < P.W && gy >= 0 && gy < P.H) {
                const long ci = (long)gy * P.W + gx;
                const unsigned f = P.fl[ci];
                if (f & kSfsInner) {
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const T* coef = t == 0 ? P.cR : P.cC;
#pragma unroll
                        for (int m = 0; m < 5; ++m) {
                            const T am = coef[(long)m * N + ci];
                            co[t * 5 + m] = am;
                            y[t] += am * sp[ly + 1 + sfsDy(t, m)][lx + 1 + sfsDx(t, m)];
                        }
                    }
                }
                if (f & kSfsFlat) {   // the 15 coefficients need 3 + 3 divisions, not 15 x 2 (sfsCoefS divides per call)
                    const T ax[3] = { sAx[lx], sAx[lx + 1], sAx[lx + 2] }, ay[3] = { sAy[ly], sAy[ly + 1], sAy[ly + 2] };   // columns gx-1..gx+1 (gx = x0 - 1 + lx), rows gy-1..gy+1
#pragma unroll
                    for (int m = 0; m < 5; ++m) {
                        const T pm = sp[ly + 1 + sfsDy(2, m)][lx + 1 + sfsDx(2, m)];
                        y[2] += sfsCoefFromA(P, ax[1 + sfsDx(2, m)], m) * pm;
                        y[3] += sfsCoefFromA(P, ay[1 + sfsDy(2, m)], m) * pm;
                        y[4] += sfsCoefFromA(P, T(1), m) * pm;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) sy[k][ly][lx] = y[k];
#pragma unroll
            for (int k = 0; k < 10; ++k) sc[k][ly][lx] = co[k];
        }
        __syncthreads();
        const int px = x0 + tx, py = y0 + ty;
        if (px < P.W && py < P.H) {
            const long u = (long)py * P.W + px;
            T o = T(0);
            if (P.fl[u] & kSfsIncluded) {
                const T xi = sp[ty + 2][tx + 2];
                o = P.wp * (P.wp * xi);                                                       // r_p
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int k = 0; k < 5; ++k) {                                             // u is dependency k of the row centred at u - offset_k
                        const int cy = ty + 1 - sfsDy(t, k), cx = tx + 1 - sfsDx(t, k);
                        o += sc[t * 5 + k][cy][cx] * sy[t][cy][cx];
                    }
                const T axu = sAx[tx + 2], ayu = sAy[ty + 2];   // dependency k of the row centred at u - offset_k is u itself: one factor pair per pixel
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int cy = ty + 1 - sfsDy(2, k), cx = tx + 1 - sfsDx(2, k);
                    o += sfsCoefFromA(P, axu, k) * sy[2][cy][cx];
                    o += sfsCoefFromA(P, ayu, k) * sy[3][cy][cx];
                    o += sfsCoefFromA(P, T(1), k) * sy[4][cy][cx];
                }
                if (LM) o += CtC[u] * xi;
                acc += xi * o;
            }
            out[u] = o;
        }
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
        tilesX = (P.W + kSfsTW - 1) / kSfsTW; numTiles = tilesX * ((P.H + kSfsTH - 1) / kSfsTH);
        int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_sfsApplyJTJTiled<T, true>, kBlock, 0);
        if (occ < 1) occ = 1;
        gridTiled = (int)std::min<long>(std::min<long>(numTiles, (long)sms * occ), kMaxPartials);
        tiled = !getenv("OPTB_SFS_NO_TILED");   // the one-pass gather (A/B measurements)
    }
    int tilesX = 0, numTiles = 0, gridTiled = 1; bool tiled = true;
    ~SfsOps() override { cudaFree(P.B); cudaFree(P.dB); cudaFree(P.cR); cudaFree(P.cC); cudaFree(P.fl); }
    const char* name() const override { return "shape_from_shading"; }
    bool capturable() const override { return true; }
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
        if (tiled) {
            if (CtC) k_sfsApplyJTJTiled<T, true><<<gridTiled, kBlock, 0, c.stream>>>(P, c.ws, dotSlot, x, out, CtC, tilesX, numTiles);
            else k_sfsApplyJTJTiled<T, false><<<gridTiled, kBlock, 0, c.stream>>>(P, c.ws, dotSlot, x, out, nullptr, tilesX, numTiles);
            return;
        }
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
