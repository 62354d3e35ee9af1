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
// Two further kernels work on tiles through the host/device functions of sfs_tile.cuh (run on the CPU by tests/test_sfs_host_parity.py):
//   k_sfsInitTiled      PCGInit1 + PCGComputeCtC + PCGFinalizeDiagonal (+ the Q0 / flag scalars) in ONE pass: the LM iteration's head
//                       was five launches (38 + 17 + 17 us + two one-thread kernels at 640x480)
//   k_sfsLoopResident   the PCG loop up to the first residual reset as ONE cooperative kernel, the tile's direction, coefficients and
//                       flags in shared memory, delta / r / b / CtC / z / Ap in registers (640x480: 140 tiles of 64x35 on 148 SMs)
#include <algorithm>
#include <type_traits>
#include "solver.h"
#include "sfs_device.cuh"
#include "sfs_tile.cuh"

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
// precompute pass 2 + computeCost in one launch: the cost of the rows centred at a pixel needs exactly what pass 2 has at hand (pass 1 of
// the pixel and of its right / lower neighbours)
template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsBuildRowsCost(SfsData<T> P, ReduceWorkspace ws, int slot) {
    const long N = (long)P.W * P.H;
    T acc = T(0);
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < N; i += (long)gridDim.x * kBlock) {
        const int x = (int)(i % P.W), y = (int)(i / P.W);
        sfsBuildRows(P, x, y);
        acc += sfsCostAt<T, false>(P, x, y, nullptr);
    }
    gridReduceTo<kBlock>((double)acc, ws, slot);
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
// ------------------------------------------------------------------------------------ precompute + computeCost, tiled
// kernels.precompute (:607-614) and computeCost (:580-592) in one pass over 32x8 tiles (sfsPreCell / sfsPrePixel in sfs_tile.cuh): dB never
// goes through global memory and the three launches of the per-pixel path (k_sfsPrecomputeB, k_sfsBuildRows, k_sfsCost) become one.
template <class T>
__global__ void __launch_bounds__(kBlock) k_sfsPrecomputeTiled(SfsData<T> P, ReduceWorkspace ws, int slot, int tilesX, int numTiles, int gate) {
    __shared__ __align__(16) unsigned char planes[sfsPrePlaneBytes<T>(kSfsTW, kSfsTH)];
    // LM decision on the device (LaunchCtx::gate): the same answer in every thread of the grid, before anything is touched
    if (gate && (ws.scalars[kLmGate] != 0.0 || (gate == 2 && ws.scalars[kLmReject] == 0.0))) return;
    const int tid = threadIdx.x, tx = tid % kSfsTW, ty = tid / kSfsTW;
    T acc = T(0);
    for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const SfsPreView<T> V = sfsCarvePre<T>(planes, (tile % tilesX) * kSfsTW, (tile / tilesX) * kSfsTH, kSfsTW, kSfsTH);
        __syncthreads();   // the previous tile's pixels are done with the planes
        for (int c = tid; c < V.bcells(); c += kBlock) sfsPreCell(P, V, c);
        __syncthreads();
        if (V.x0 + tx < P.W && V.y0 + ty < P.H) acc += sfsPrePixel(P, V, ty, tx);
    }
    gridReduceTo<kBlock>((double)acc, ws, slot);
}

// ------------------------------------------------------------------------------------ PCGInit1 (+ LM head), tiled
// One pass over 32x8 tiles: the residual of every row centred in tile + 1-pixel ring is formed ONCE (the per-pixel k_sfsInit forms each
// 5 / 15 times), then every pixel gathers -J^T F and diag J^T J from shared memory and finishes PCGInit1; LM: PCGSaveSSq (first
// iteration), PCGComputeCtC and PCGFinalizeDiagonal are applied to the pixel right away and Q0 = q = 0 and the early-out flag are reset,
// i.e. k_sfsInit + k_sfsCtC + k_lmFinalizeDiagonal + k_copyScalar + k_setScalar of the launch-by-launch path (solver.cu).
template <class T, bool LM>
__global__ void __launch_bounds__(kBlock) k_sfsInitTiled(SfsData<T> P, ReduceWorkspace ws, int rzSlot, SfsInitOut<T> O, int usePre, int saveSSq, T minDiag, T maxDiag,
                                                         int tilesX, int numTiles) {
    __shared__ __align__(16) unsigned char planes[sfsTilePlaneBytes<T>(kSfsTW, kSfsTH)];
    const int tid = threadIdx.x, tx = tid % kSfsTW, ty = tid / kSfsTW;
    const T radius = LM ? (T)ws.scalars[kRadius] : T(1);
    if (LM && blockIdx.x == 0 && tid == 0) { ws.scalars[kQ] = 0.0; ws.scalars[kQPrev] = 0.0; ws.scalars[kFlags] = 0.0; }   // delta = 0 => q = 0 (:631-664), Q0 = q, no early-out yet
    T acc = T(0);
    for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const SfsTileView<T> V = sfsCarveTile<T>(planes, (tile % tilesX) * kSfsTW, (tile / tilesX) * kSfsTH, kSfsTW, kSfsTH);
        __syncthreads();   // the previous tile's gathers are done with the planes
        for (int i = tid; i < V.PW() + V.PH(); i += kBlock) sfsTileFactors(P, V, i);
        for (int c = tid; c < V.cells(); c += kBlock) sfsTileStageCell(P, V, c);
        __syncthreads();
        for (int c = tid; c < V.cells(); c += kBlock) sfsTileRowResiduals(P, V, c);
        __syncthreads();
        const int px = V.x0 + tx, py = V.y0 + ty;
        if (px < P.W && py < P.H) {
            const long u = (long)py * P.W + px;
            const bool inc = (V.sfl[(ty + 1) * V.RW() + tx + 1] & kSfsIncluded) != 0;
            T jtf = T(0), dg = T(0);
            if (inc) { jtf = sfsTileGather(P, V, ty, tx, P.wp * sfsFP(P, u)); if (LM || usePre) dg = sfsTileGatherDiag(P, V, ty, tx); }
            acc += sfsInitPixel<T, LM>(inc, jtf, dg, u, O, usePre, saveSSq, radius, minDiag, maxDiag);
        }
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}

// ------------------------------------------------------------------------------------ whole PCG loop, state on chip
// L iterations of PCGStep1/2/3 (:1056-1102) from the state PCGInit1 (+ FinalizeDiagonal) left, as one cooperative kernel: one CTA per
// tw x th tile (sfs_tile.cuh).  Per iteration only the z of pixels within two of a tile border crosses CTAs (as tagged 8-byte words the
// neighbours poll at the top of the next iteration: no fence, no barrier) plus the two all-reduces (p.Ap; then r.z and q in one round).  LM: zeta = (l+1)(Q1-Q0)/Q1 < q_tolerance ends the loop in every CTA at once
// (they all hold the same reduced q); the flag word tells the per-step kernels of a following residual-reset iteration to do nothing.
// The reduction sequence number lives in device memory (seqWord) so that the launch is replayable from a captured graph.
// Tail (LM): computeModelCost of the delta found (one more tile-edge exchange of delta, the rows applied to it, F + J delta summed) and
// the update of the unknowns, so that a nonlinear iteration is head + this kernel + precompute/cost.
template <class T> struct SfsLoopHead { int on, saveSSq; T minDiag, maxDiag; T* SSq; T* pre; T* CtC; T* b; };   // the LM head inside the loop kernel (on = 0: the vectors come from memory)
template <class T, int KMAX, int NT>
__global__ void __launch_bounds__(NT, 1) k_sfsLoopResident(SfsData<T> P, ReduceWorkspace ws, unsigned long long* llWords, unsigned long long* arrivals, unsigned long long* seqWord,
                                                           unsigned long long* zLL, unsigned long long* dLL,
                                                             int L, int tw, int th, int tilesX, int usePre, int lm, int resetPeriod, int withModelCost, T qTol,
                                                             T* delta, T* r, T* z, T* p, T* Ap, const T* pre, const T* b, const T* CtC, T* xUpdate, T* xSaved, SfsLoopHead<T> head, int gate) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int tid = threadIdx.x;
    // LM decision on the device: an iteration enqueued behind the one that ended the solve does nothing -- every CTA reads the same word
    // (written by k_lmDecide, an earlier kernel on this stream) and leaves before the first all-reduce, so the sequence word stays put
    if (gate && ws.scalars[kLmGate] != 0.0) return;
    const SfsTileView<T> V = sfsCarveTile<T>(smemRaw, ((int)blockIdx.x % tilesX) * tw, ((int)blockIdx.x / tilesX) * th, tw, th, true);
    SfsLoopRegs<T, KMAX> R;
    unsigned long long k = ldVolatileU64(seqWord);   // written by the previous launch on this stream; every CTA reads it before the first all-reduce can complete
    double dRz = 0.0, dRzPrev = 0.0, dPAp = 0.0, dQ = 0.0, unused = 0.0;
    T Q0 = T(0);
    int done = 0, broke = -1;
    for (int i = tid; i < V.PW() + V.PH(); i += NT) sfsTileFactors(P, V, i);
    for (int c = tid; c < V.cells(); c += NT) sfsTileStageCell(P, V, c);
    if (head.on) {
        // the nonlinear iteration's head on the tile (what k_sfsInitTiled does over the image): residuals of the staged rows, then
        // -J^T F, diag J^T J and the LM diagonal / preconditioner / first direction per pixel, straight into the loop state
        for (int c = tid; c < (V.PW()) * (V.PH()); c += NT) { V.sp[c] = T(0); V.sd[c] = T(0); }
        __syncthreads();
        for (int c = tid; c < V.cells(); c += NT) sfsTileRowResiduals(P, V, c);
        __syncthreads();
        const T a0 = sfsLoopHead<T, KMAX>(P, V, R, tid, NT, usePre, head.saveSSq, (T)ws.scalars[kRadius], head.minDiag, head.maxDiag, head.SSq, zLL, sfsEdgeTag(k, kSfsEdgeFirstDirection));
        sfsLoopFetchRing<T>(P, V, tid, NT, V.sp, zLL, sfsEdgeTag(k, kSfsEdgeFirstDirection));   // the neighbours' first direction on the ring
        gridAllReduce2LL<NT, false, false>((double)a0, 0.0, llWords, arrivals, ++k, ws.comm, 0, dRz, unused);   // r . p (its barriers also order the planes)
    } else {
        sfsLoopLoad<T, KMAX>(P, V, R, tid, NT, delta, r, b, CtC, pre, p, usePre, lm);
        dRz = ws.scalars[kRz0];
        Q0 = lm ? (T)ws.scalars[kQPrev] : T(0);
        __syncthreads();
    }
    for (int l = 0; l < L; ++l) {
        if (l > 0) {   // PCGStep3 of iteration l - 1
            T beta = T(0);
            { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
            sfsLoopUpdateDirection<T, KMAX>(P, V, R, tid, NT, beta, zLL, sfsEdgeTag(k - 1, kSfsEdgeZ));   // the z published before the last all-reduce
            __syncthreads();
        }
        for (int c = tid; c < V.cells(); c += NT) sfsTileRowsTimesVec(P, V, c, V.sp);
        __syncthreads();
        const T acc = sfsLoopApply<T, KMAX>(P, V, R, tid, NT);
        gridAllReduce2LL<NT, false, false>((double)acc, 0.0, llWords, arrivals, ++k, ws.comm, 0, dPAp, unused);
        T alpha = T(0);
        { const T num = (T)dRz, den = (T)dPAp; if (den > T(0)) alpha = num / den; }
        T accRz = T(0), accQ = T(0);
        if (lm && resetPeriod > 0 && ((l + 1) % resetPeriod) == 0) {   // residual reset (:1077-1083): r = b - A delta instead of r -= alpha A p
            sfsLoopResetFirstHalf<T, KMAX>(P, V, R, tid, NT, alpha, dLL, sfsEdgeTag(k, kSfsEdgeDeltaReset));
            sfsLoopFetchRing<T>(P, V, tid, NT, V.sd, dLL, sfsEdgeTag(k, kSfsEdgeDeltaReset));   // waits word by word for the neighbours' delta
            __syncthreads();
            for (int c = tid; c < V.cells(); c += NT) sfsTileRowsTimesVec(P, V, c, V.sd);
            __syncthreads();
            sfsLoopResetSecondHalf<T, KMAX>(P, V, R, tid, NT, usePre, zLL, sfsEdgeTag(k, kSfsEdgeZ), accRz, accQ);
        } else {
            sfsLoopStep2<T, KMAX>(P, V, R, tid, NT, alpha, usePre, zLL, sfsEdgeTag(k, kSfsEdgeZ), accRz, accQ);
        }
        dRzPrev = dRz;
        gridAllReduce2LL<NT, false, false>((double)accRz, (double)accQ, llWords, arrivals, ++k, ws.comm, 0, dRz, dQ);
        done = l + 1;
        if (lm) {
            const T Q1 = (T)dQ;
            const T zeta = T(l + 1) * (Q1 - Q0) / Q1;
            Q0 = Q1;
            if (zeta < qTol) { broke = l; break; }   // the reference updates p once more before it leaves (the epilogue does)
        }
    }
    if (withModelCost) {   // computeModelCost (:666-678) for the delta just found: its rows need delta on a 1-pixel ring
        double u1, dModel;
        sfsLoopPublishDelta<T, KMAX>(P, V, R, tid, NT, dLL, sfsEdgeTag(k, kSfsEdgeDeltaModel));
        sfsLoopFetchRing<T>(P, V, tid, NT, V.sd, dLL, sfsEdgeTag(k, kSfsEdgeDeltaModel));
        __syncthreads();
        for (int c = tid; c < V.cells(); c += NT) sfsTileRowsTimesVec(P, V, c, V.sd);
        __syncthreads();
        const T e = sfsLoopModelCost<T, KMAX>(P, V, R, tid, NT);
        gridAllReduce2LL<NT, false, false>((double)e, 0.0, llWords, arrivals, ++k, ws.comm, 0, dModel, u1);
        if (blockIdx.x == 0 && tid == 0) ws.scalars[kModelCost] = dModel;
    }
    T beta = T(0);
    { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
    // xUpdate: savePreviousUnknowns + PCGLinearUpdate (:1108-1113, :552-557) for the tile as well -- nobody reads X any more (the model cost's
    // rows did, across tiles, but every CTA is past its all-reduce)
    sfsLoopStore<T, KMAX>(P, V, R, tid, NT, done > 0, beta, delta, r, z, p, Ap, xUpdate, xSaved, head.on ? head.pre : nullptr, head.CtC, head.b);
    if (blockIdx.x == 0 && tid == 0 && done > 0) {   // leave the scalars as the kernel-per-step loop does (iteration l reads rz[l&1], writes rz[(l+1)&1])
        ws.scalars[kPAp] = dPAp; ws.scalars[done & 1] = dRz; ws.scalars[(done + 1) & 1] = dRzPrev;
        if (lm) { ws.scalars[kQ] = dQ; ws.scalars[kQPrev] = (double)Q0; if (broke >= 0) ws.scalars[kFlags] = 1.0 + broke; else if (head.on) ws.scalars[kFlags] = 0.0; }
        *seqWord = k;
    }
}
__global__ void k_sfsCoopProbe(int* out) { if (out) *out = 1; }
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
        { int o2 = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, (const void*)k_sfsInitTiled<T, true>, kBlock, 0); if (o2 < 1) o2 = 1;
          gridInit = (int)std::min<long>(std::min<long>(numTiles, (long)sms * o2), kMaxPartials); }
        fusedHead = tiled && !getenv("OPTB_SFS_NO_FUSED_INIT");
        fuseCost = !getenv("OPTB_SFS_NO_FUSED_COST");
        fusePre = fuseCost && !getenv("OPTB_SFS_NO_FUSED_PRECOMPUTE");
        { int o3 = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, (const void*)k_sfsPrecomputeTiled<T>, kBlock, 0); if (o3 < 1) o3 = 1;
          gridPre = (int)std::min<long>(std::min<long>(numTiles, (long)sms * o3), kMaxPartials); }
        fuseModel = !getenv("OPTB_SFS_NO_FUSED_MODEL");
        fuseUpdate = !getenv("OPTB_SFS_NO_FUSED_UPDATE");
        fuseHeadInLoop = !getenv("OPTB_SFS_NO_HEAD_IN_LOOP");
        if (tiled && !getenv("OPTB_SFS_NO_RESIDENT")) setUpResidentLoop(dev, sms);
    }
    int tilesX = 0, numTiles = 0, gridTiled = 1, gridInit = 1; bool tiled = true, fusedHead = true;
    // ---- whole-loop kernel: tile shape, shared memory, reduction words (all fixed for the life of the plan)
    static constexpr int kLoopK = 3;              // pixels per thread
    int loopTW = 0, loopTH = 0, loopTilesX = 0, loopGrid = 0, loopNT = 768; size_t loopSmem = 0;
    const void* loopFn() const {
        return loopNT == 768 ? (const void*)k_sfsLoopResident<T, kLoopK, 768> : (loopNT == 832 ? (const void*)k_sfsLoopResident<T, kLoopK, 832> : (const void*)k_sfsLoopResident<T, kLoopK, 1024>);
    }
    unsigned long long* loopWords = nullptr;      // [2 parities][kMaxLLCtas + 1][4] reduction words, arrival counter, sequence word
    unsigned long long* loopEdge = nullptr;       // tile-edge exchange words: z, then delta
    void setUpResidentLoop(int dev, int sms) {
        int coop = 0, maxSmem = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (!coop) return;
        // threads per CTA: 1024 leaves the third pixel round of a 64x35 tile 19 % full; 832 and 768 fill their three rounds better with fewer
        // warps to hide latency (A/B on hardware: OPTB_SFS_LOOP_NT)
        loopNT = 768;   // measured at 640x480: 15.4 us per PCG step against 15.6 (832) and 16.2 (1024)
        if (const char* e = getenv("OPTB_SFS_LOOP_NT")) { const int v = atoi(e); if (v == 768 || v == 832 || v == 1024) loopNT = v; }
        const void* fn = loopFn();
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, fn) != cudaSuccess) { cudaGetLastError(); return; }
        // one tile per SM at most; tile widths in multiples of 32 (warp = one tile row segment), the squarest shape that fits wins:
        // its ring (what crosses CTAs every step) and its (tile + ring) planes are the smallest
        long best = -1;
        for (int tw = 32; tw <= (P.W + 31) / 32 * 32 && tw <= 1024; tw += 32) {
            const int tx = (P.W + tw - 1) / tw;
            if (tx > sms || tx > kMaxLLCtas) continue;
            int th = (P.H + sms / tx - 1) / (sms / tx);
            if (th < 4) th = 4;
            if ((long)tw * th > (long)kLoopK * loopNT) continue;
            const size_t bytes = sfsTilePlaneBytes<T>(tw, th, true);
            if (bytes + fa.sharedSizeBytes > (size_t)maxSmem) continue;
            const long cost = (long)(tw + 2) * (th + 2) + (long)(tw + 4) * (th + 4);
            if (best < 0 || cost < best) { best = cost; loopTW = tw; loopTH = th; loopTilesX = tx; loopGrid = tx * ((P.H + th - 1) / th); loopSmem = bytes; }
        }
        if (best < 0) { loopTW = 0; return; }
        int occ = 0;
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)loopSmem) != cudaSuccess
            || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, loopNT, loopSmem) != cudaSuccess || occ < 1 || loopGrid > sms * occ || loopGrid > kMaxLLCtas) {
            cudaGetLastError(); loopTW = 0; return;
        }
        // a cooperative launch inside a stream capture must work: the loop kernel is a node of the plan's PCG-loop / iteration graphs
        {
            cudaStream_t s = nullptr; cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr; bool ok = false;
            int* none = nullptr; void* args[] = { (void*)&none };
            if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess) {
                if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const cudaError_t le = cudaLaunchCooperativeKernel((const void*)k_sfsCoopProbe, dim3(1), dim3(32), args, 0, s);
                    const cudaError_t ee = cudaStreamEndCapture(s, &g);
                    if (le == cudaSuccess && ee == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess)
                        ok = cudaGraphLaunch(ge, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
                }
                if (ge) cudaGraphExecDestroy(ge);
                if (g) cudaGraphDestroy(g);
                cudaStreamDestroy(s);
            }
            cudaGetLastError();
            if (!ok) { loopTW = 0; return; }
        }
        constexpr size_t words = 2 * 4 * (kMaxLLCtas + 1) + 2;
        OPT_CUDA_CHECK(cudaMalloc((void**)&loopWords, sizeof(unsigned long long) * words));
        OPT_CUDA_CHECK(cudaMemset(loopWords, 0, sizeof(unsigned long long) * words));
        const size_t edgeWords = (size_t)N * LLWords<T>::N;   // tile-edge words of z and of delta, indexed like the image (tag 0 = never written)
        OPT_CUDA_CHECK(cudaMalloc((void**)&loopEdge, sizeof(unsigned long long) * 2 * edgeWords));
        OPT_CUDA_CHECK(cudaMemset(loopEdge, 0, sizeof(unsigned long long) * 2 * edgeWords));
    }
    ~SfsOps() override { cudaFree(P.B); cudaFree(P.dB); cudaFree(P.cR); cudaFree(P.cC); cudaFree(P.fl); cudaFree(loopWords); cudaFree(loopEdge); }
    const char* name() const override { return "shape_from_shading"; }
    bool capturable() const override { return true; }
    unsigned long long captureKey() const override { const int shape[] = { loopTW, loopTH, loopNT, (int)fusedHead, (int)tiled, (int)fuseCost, (int)fusePre, (int)fuseModel, (int)fuseUpdate, (int)fuseHeadInLoop }; return EnergyOps<T>::hashBytes(shape, sizeof(shape), EnergyOps<T>::hashValue(P)); }
    void bind(void** q) override {   // positions: shape_from_shading.t:5-28 (all scalar Params are float, cast to opt_float on use, o.t:1699)
        auto f = [&](int k) { return (T) * (const float*)q[k]; };
        costFresh = false;
        P.wp = (T)sqrt((double)f(0)); P.ws = (T)sqrt((double)f(1)); P.wg = (T)sqrt((double)f(2));
        if (std::is_same<T, float>::value) { P.wp = (T)sqrtf(*(const float*)q[0]); P.ws = (T)sqrtf(*(const float*)q[1]); P.wg = (T)sqrtf(*(const float*)q[2]); }
        P.fx = f(3); P.fy = f(4); P.ux = f(5); P.uy = f(6);
        for (int i = 1; i <= 9; ++i) P.L[i] = f(6 + i);
        this->unknowns[0].user = q[16];
        P.X = (const T*)q[16]; P.D = (const T*)q[17]; P.Im = (const T*)q[18]; P.eR = (const unsigned char*)q[19]; P.eC = (const unsigned char*)q[20];
    }
    // precompute also leaves the cost of the X it linearised at in scalars[kCost]; the cost() that follows it (every call site in
    // solver.cu pairs them) is then free
    bool fuseCost = true, fusePre = true, costFresh = false; int gridPre = 1;
    void precompute(const LaunchCtx& c) override {
        if (fusePre) { k_sfsPrecomputeTiled<T><<<gridPre, kBlock, 0, c.stream>>>(P, c.ws, kCost, tilesX, numTiles, c.gate); costFresh = true; return; }
        if (c.gate) EnergyOps<T>::unsupported("a gated precompute without the fused precompute + cost pass");
        k_sfsPrecomputeB<T><<<grid, kBlock, 0, c.stream>>>(P);
        if (fuseCost) { k_sfsBuildRowsCost<T><<<grid, kBlock, 0, c.stream>>>(P, c.ws, kCost); costFresh = true; }
        else k_sfsBuildRows<T><<<grid, kBlock, 0, c.stream>>>(P);
    }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        if (fusedHead) {
            const SfsInitOut<T> O{ delta, r, pre, p, nullptr, nullptr, nullptr };
            k_sfsInitTiled<T, false><<<gridInit, kBlock, 0, c.stream>>>(P, c.ws, rzSlot, O, this->usePreconditioner ? 1 : 0, 0, T(0), T(0), tilesX, numTiles);
            return;
        }
        k_sfsInit<T><<<grid, kBlock, 0, c.stream>>>(P, this->usePreconditioner ? 1 : 0, c.ws, rzSlot, delta, r, pre, p);
    }
    bool pcgInitLM(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, T* CtC, T* SSq, T* b, int rzSlot, T minDiag, T maxDiag, bool saveSSq) override {
        if (!fusedHead) return false;
        const SfsInitOut<T> O{ delta, r, pre, p, CtC, SSq, b };
        k_sfsInitTiled<T, true><<<gridInit, kBlock, 0, c.stream>>>(P, c.ws, rzSlot, O, this->usePreconditioner ? 1 : 0, saveSSq ? 1 : 0, minDiag, maxDiag, tilesX, numTiles);
        return true;
    }
    bool fuseModel = true, fuseUpdate = true;
    bool fuseHeadInLoop = true;
    bool pcgLoopTakesHead(int L, bool lm) override { return loopTW && L >= 2 && lm && fuseHeadInLoop && fusedHead; }
    // head + loop + model cost + update in the loop kernel, precompute + cost in one pass: both kernels test LaunchCtx::gate
    bool deviceLmCapable() const override { return loopTW && fusedHead && fuseHeadInLoop && fuseModel && fuseUpdate && fusePre; }
    int pcgLoopOnChip(const LaunchCtx& c, int L, bool lm, int resetPeriod, T qTol, T* delta, T* r, T* z, T* p, T* Ap, const T* pre, const T* b, const T* CtC,
                      bool* modelCostDone, T* prevX, bool* updateDone, const typename EnergyOps<T>::LoopHeadArgs* headArgs) override {
        if (!loopTW || L < 2) { if (headArgs) EnergyOps<T>::unsupported("pcgLoopOnChip with the head after pcgLoopTakesHead said yes"); return 0; }
        SfsLoopHead<T> head{};
        if (headArgs) { head.on = 1; head.saveSSq = headArgs->saveSSq ? 1 : 0; head.minDiag = headArgs->minDiag; head.maxDiag = headArgs->maxDiag; head.SSq = headArgs->SSq;
                        head.pre = const_cast<T*>(pre); head.CtC = const_cast<T*>(CtC); head.b = const_cast<T*>(b); }
        int withModel = (lm && fuseModel && modelCostDone) ? 1 : 0;
        if (modelCostDone) *modelCostDone = withModel != 0;
        // the one unknown array of this energy, offset 0 in the plan's vectors.  LM: only together with the model cost, which must see the old X
        T* xUpdate = (updateDone && fuseUpdate && (!lm || (prevX && withModel))) ? const_cast<T*>(P.X) : nullptr;
        T* xSaved = lm ? prevX : nullptr;
        if (updateDone) *updateDone = xUpdate != nullptr;
        ReduceWorkspace ws = c.ws;
        unsigned long long* arrivals = loopWords + 2 * 4 * (kMaxLLCtas + 1); unsigned long long* seqWord = arrivals + 1;
        int usePre = this->usePreconditioner ? 1 : 0, lmi = lm ? 1 : 0, gate = c.gate;
        unsigned long long* zLL = loopEdge; unsigned long long* dLL = loopEdge + (size_t)N * LLWords<T>::N;
        void* args[] = { (void*)&P, (void*)&ws, (void*)&loopWords, (void*)&arrivals, (void*)&seqWord, (void*)&zLL, (void*)&dLL, (void*)&L, (void*)&loopTW, (void*)&loopTH, (void*)&loopTilesX, (void*)&usePre, (void*)&lmi,
                         (void*)&resetPeriod, (void*)&withModel, (void*)&qTol, (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&Ap, (void*)&pre, (void*)&b, (void*)&CtC, (void*)&xUpdate, (void*)&xSaved, (void*)&head, (void*)&gate };
        OPT_CUDA_CHECK(cudaLaunchCooperativeKernel(loopFn(), dim3((unsigned)loopGrid), dim3((unsigned)loopNT), args, loopSmem, c.stream));
        return L;
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
    void cost(const LaunchCtx& c) override {
        if (costFresh) { costFresh = false; return; }
        k_sfsCost<T, false><<<grid, kBlock, 0, c.stream>>>(P, c.ws, kCost, nullptr);
    }
    void modelCost(const LaunchCtx& c, const T* delta) override { k_sfsCost<T, true><<<grid, kBlock, 0, c.stream>>>(P, c.ws, kModelCost, delta); }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override { k_sfsCtC<T><<<grid, kBlock, 0, c.stream>>>(P, CtC, invRadius); }
};

template <class T> EnergyOps<T>* makeSfs(const unsigned* dims) { return new SfsOps<T>(dims); }
template EnergyOps<float>* makeSfs<float>(const unsigned*);
template EnergyOps<double>* makeSfs<double>(const unsigned*);

}  // namespace optb
