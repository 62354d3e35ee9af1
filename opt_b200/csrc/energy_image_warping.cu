// energy_image_warping.cu -- hand-written sm_100a kernels for the image_warping (2-D ARAP) energy.
//
// Energy (reference examples/image_warping/image_warping.t:1-23), per pixel i with 4-neighbours j:
//     r_reg(i,j) = w_reg * [ (O_i - O_j) - R(a_i) (U_i - U_j) ]     valid iff j in bounds, Mask_j==0, Mask_i==0
//     r_fit(i)   = w_fit * (O_i - C_i)                              valid iff C_i.x>=0 and C_i.y>=0
//     Exclude(Mask_i != 0);  UsePreconditioner(true);  unknowns O (2 ch) and a (1 ch).
// What the reference generates symbolically for this spec (o.t createjtjcentered :2029-2089,
// createjtfcentered :2129-2172, createcost :2375-2385, computeCtCcentered :2255-2287,
// createmodelcost :2174-2199) is written out here in closed form.  With v_ij the (symmetric) validity
// of the pair, d_ij = U_i - U_j, R' = dR/da and e_i = R'(a_i) d_ij:
//     (J^T J p)_O(i) = w^2 sum_j v_ij [ 2 (pO_i - pO_j) - e_i pa_i - e_j pa_j ] + w_fit^2 vfit_i pO_i
//     (J^T J p)_a(i) = w^2 sum_j v_ij [ |e_i|^2 pa_i - e_i . (pO_i - pO_j) ]
//     (J^T F)_O(i)   = w sum_j v_ij [ r_ij - r_ji ] + w_fit r_fit(i)
//     (J^T F)_a(i)   = -w sum_j v_ij e_i . r_ij
//     diag_O(i) = 2 w^2 #valid + w_fit^2 vfit ,  diag_a(i) = w^2 sum_j v_ij |e_i|^2
// (the sum over j collects both the residual centred at i and the one centred at j that reads O_i,
// o.t residualsincludingX00 :1996-2006).
//
// B200 mapping: the unknowns are fixed during the inner PCG loop, so PCGInit1 packs everything
// JtJ needs into plan-owned arrays -- geom = (cos a, sin a, Ux, Uy) as one 16-byte vector per pixel
// and a flag byte (pair validity x4, fit validity, active) -- and the inner loop never touches the
// caller's buffers, never evaluates sin/cos (the reference's generated Step1 evaluates ~5 per
// pixel), and reads 7 floats + 1 byte per pixel instead of the spec's 9 + mask.  HBM-bound stencil:
// one thread per pixel over 32x8 tiles, neighbours through L1/L2, persistent grid sized to the SMs.
#include <algorithm>
#include <cstdint>
#include <map>
#include <type_traits>
#include <cuda.h>
#include "solver.h"

namespace optb {

template <class T> struct alignas(sizeof(T) * 2) Vec2 { T x, y; };
template <class T> struct alignas(16) Vec4 { T x, y, z, w; };

template <class T>
struct IWData {
    int W, H, tilesX, numTiles;
    const Vec2<T>* Offset; const T* Angle; const Vec2<T>* UrShape; const Vec2<T>* Constraints; const T* Mask;
    T w, wf;                 // w_regSqrt, w_fitSqrt (float Params cast to opt_float on use, o.t:1699)
    long offA;               // offset of the Angle block inside the flat PCG vectors
    Vec4<T>* geom; uint8_t* flags;
    int y0, y1;              // owned local rows (multi-GPU slab; [0,H) on a single GPU).  Rows outside are halo:
                             // read as neighbours, never written as results, flags = 0
};

constexpr int kTileW = 32, kTileH = 8;
__device__ __forceinline__ void sincosT(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincosT(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float guardedInvert(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }    // CERES variant, solverGPUGaussNewton.t:323-332
__device__ __forceinline__ double guardedInvert(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }

#define IW_TILE_LOOP(D)                                                                   \
    for (int tile = blockIdx.x; tile < (D).numTiles; tile += gridDim.x)                   \
        for (int _once = 0, x = (tile % (D).tilesX) * kTileW + (threadIdx.x & 31),        \
                 y = (tile / (D).tilesX) * kTileH + (threadIdx.x >> 5);                   \
             _once < 1; ++_once)                                                          \
            if (x < (D).W && y < (D).H)

// neighbour k of the spec's Stencil {{1,0},{-1,0},{0,1},{0,-1}}
__device__ __forceinline__ bool neighbour(int k, int x, int y, int W, int H, long* j) {
    const int nx = x + (k == 0 ? 1 : (k == 1 ? -1 : 0));
    const int ny = y + (k == 2 ? 1 : (k == 3 ? -1 : 0));
    *j = (long)ny * W + nx;
    return nx >= 0 && nx < W && ny >= 0 && ny < H;
}

// ------------------------------------------------------------------------------------ PCGInit1
// solverGPUGaussNewton.t:361-397 specialised to this energy; also packs geom/flags.
template <class T>
__global__ void __launch_bounds__(kBlock) k_iwInit(IWData<T> D, ReduceWorkspace ws, int rzSlot, T* __restrict__ delta, T* __restrict__ r,
                                                   T* __restrict__ pre, T* __restrict__ p) {
    T acc = T(0);
    IW_TILE_LOOP(D) {
        const long i = (long)y * D.W + x;
        const Vec2<T> Ui = D.UrShape[i];
        const T ai = D.Angle[i];
        T s, c; sincosT(ai, &s, &c);
        D.geom[i] = Vec4<T>{ c, s, Ui.x, Ui.y };
        uint8_t fl = 0;
        Vec2<T> rO{ T(0), T(0) }, pO{ T(0), T(0) }, preO{ T(0), T(0) };
        T rA = T(0), pA = T(0), preA = T(0);
        if (y >= D.y0 && y < D.y1 && D.Mask[i] == T(0)) {
            const Vec2<T> Oi = D.Offset[i];
            T gx = T(0), gy = T(0), gA = T(0), dO = T(0), dA = T(0);
            const T w = D.w, w2 = D.w * D.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                long j;
                if (neighbour(k, x, y, D.W, D.H, &j) && D.Mask[j] == T(0)) {
                    fl |= (uint8_t)(1u << k);
                    const Vec2<T> Oj = D.Offset[j], Uj = D.UrShape[j];
                    T sj, cj; sincosT(D.Angle[j], &sj, &cj);
                    const T dx = Ui.x - Uj.x, dy = Ui.y - Uj.y;
                    const T ox = Oi.x - Oj.x, oy = Oi.y - Oj.y;
                    const T rijx = w * (ox - (c * dx - s * dy)), rijy = w * (oy - (s * dx + c * dy));
                    const T rjix = w * (-ox + (cj * dx - sj * dy)), rjiy = w * (-oy + (sj * dx + cj * dy));
                    gx += w * (rijx - rjix); gy += w * (rijy - rjiy);
                    const T ex = -s * dx - c * dy, ey = c * dx - s * dy;   // e_i = R'(a_i) d
                    gA -= w * (ex * rijx + ey * rijy);
                    dO += T(2) * w2;
                    dA += w2 * (ex * ex + ey * ey);
                }
            }
            const Vec2<T> C = D.Constraints[i];
            if (C.x >= T(0) && C.y >= T(0)) {
                fl |= 16;
                gx += D.wf * (D.wf * (Oi.x - C.x)); gy += D.wf * (D.wf * (Oi.y - C.y));
                dO += D.wf * D.wf;
            }
            fl |= 32;
            rO = Vec2<T>{ -gx, -gy }; rA = -gA;
            const T io = guardedInvert(dO), ia = guardedInvert(dA);
            preO = Vec2<T>{ io, io }; preA = ia;
            pO = Vec2<T>{ io * rO.x, io * rO.y }; pA = ia * rA;
            acc += rO.x * pO.x + rO.y * pO.y + rA * pA;
        }
        else if (D.Mask[i] == T(0)) {
            // halo row (multi-GPU): no unknowns of ours here, but the pair-validity bits are still needed by the
            // owned row next to it (the tiled kernel reads the validity of a vertical pair from its upper pixel)
#pragma unroll
            for (int k = 0; k < 4; ++k) { long j; if (neighbour(k, x, y, D.W, D.H, &j) && D.Mask[j] == T(0)) fl |= (uint8_t)(1u << k); }
        }
        D.flags[i] = fl;
        reinterpret_cast<Vec2<T>*>(delta)[i] = Vec2<T>{ T(0), T(0) }; delta[D.offA + i] = T(0);
        reinterpret_cast<Vec2<T>*>(r)[i] = rO; r[D.offA + i] = rA;
        reinterpret_cast<Vec2<T>*>(pre)[i] = preO; pre[D.offA + i] = preA;
        if (y >= D.y0 && y < D.y1) { reinterpret_cast<Vec2<T>*>(p)[i] = pO; p[D.offA + i] = pA; }   // halo rows of p arrive from the neighbour
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}

// ------------------------------------------------------------------------------------ PCGStep1 (+ fused PCGStep3)
// solverGPUGaussNewton.t:421-434 (and :537-550 when FUSE3): out = (J^T J [+CtC]) p, dot = p.out.
// FUSE3: p = z + beta*pOld is formed on the fly for the centre and its neighbours; the centre value
// is written to pNew (double-buffered so neighbours in other CTAs still see pOld).
template <class T, bool FUSE3, bool LM>
__global__ void __launch_bounds__(kBlock) k_iwApplyJTJ(IWData<T> D, ReduceWorkspace ws, int dotSlot, int rzNew, int rzOld,
                                                       const T* __restrict__ pIn, const T* __restrict__ z, T* __restrict__ pOut,
                                                       T* __restrict__ out, const T* __restrict__ CtC) {
    T beta = T(0);
    if (FUSE3) {
        double drn, dro; loadScalars2(ws, rzNew, rzOld, drn, dro);
        const T rn = (T)drn, ro = (T)dro;
        if (ro > T(0)) beta = rn / ro;
    }
    const Vec2<T>* pInO = reinterpret_cast<const Vec2<T>*>(pIn);
    const T* pInA = pIn + D.offA;
    const Vec2<T>* zO = reinterpret_cast<const Vec2<T>*>(z);
    const T* zA = z + D.offA;
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    T acc = T(0);
    IW_TILE_LOOP(D) {
        const long i = (long)y * D.W + x;
        const unsigned fl = D.flags[i];
        Vec2<T> oO{ T(0), T(0) }; T oA = T(0);
        Vec2<T> pi{ T(0), T(0) }; T pai = T(0);
        if (FUSE3) {   // every pixel (halo rows included) gets its new direction; only active owned pixels get an Ap
            pi = pInO[i]; pai = pInA[i];
            const Vec2<T> zi = zO[i]; pi.x = zi.x + beta * pi.x; pi.y = zi.y + beta * pi.y; pai = zA[i] + beta * pai;
        }
        if (fl & 32u) {
            const Vec4<T> gi = D.geom[i];
            if (!FUSE3) { pi = pInO[i]; pai = pInA[i]; }
            T ax = T(0), ay = T(0), aa = T(0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (fl & (1u << k)) {
                    long j; neighbour(k, x, y, D.W, D.H, &j);
                    const Vec4<T> gj = D.geom[j];
                    Vec2<T> pj = pInO[j]; T paj = pInA[j];
                    if (FUSE3) { const Vec2<T> zj = zO[j]; pj.x = zj.x + beta * pj.x; pj.y = zj.y + beta * pj.y; paj = zA[j] + beta * paj; }
                    const T dx = gi.z - gj.z, dy = gi.w - gj.w;
                    const T eix = -gi.y * dx - gi.x * dy, eiy = gi.x * dx - gi.y * dy;
                    const T ejx = -gj.y * dx - gj.x * dy, ejy = gj.x * dx - gj.y * dy;
                    const T px = pi.x - pj.x, py = pi.y - pj.y;
                    ax += T(2) * px - eix * pai - ejx * paj;
                    ay += T(2) * py - eiy * pai - ejy * paj;
                    aa += (eix * eix + eiy * eiy) * pai - (eix * px + eiy * py);
                }
            }
            oO.x = w2 * ax; oO.y = w2 * ay; oA = w2 * aa;
            if (fl & 16u) { oO.x += wf2 * pi.x; oO.y += wf2 * pi.y; }
            if (LM) {
                const Vec2<T> cO = reinterpret_cast<const Vec2<T>*>(CtC)[i];
                oO.x += cO.x * pi.x; oO.y += cO.y * pi.y; oA += CtC[D.offA + i] * pai;
            }
            acc += pi.x * oO.x + pi.y * oO.y + pai * oA;
        }
        reinterpret_cast<Vec2<T>*>(out)[i] = oO; out[D.offA + i] = oA;
        if (FUSE3) { reinterpret_cast<Vec2<T>*>(pOut)[i] = pi; pOut[D.offA + i] = pai; }
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}

// ------------------------------------------------------------------------------------ PCGStep1, TMA-tiled (fp32)
// Same operator as k_iwApplyJTJ, different data movement and schedule.  A persistent CTA walks
// 64x16-pixel tiles; one elected thread streams each tile (+halo) of z, p, geom and the flag bytes into
// shared memory with cp.async.bulk.tensor (TMA) under an mbarrier, one tile ahead of the math, so HBM
// latency is hidden by the copy engine rather than by warp occupancy (the direct-load kernel stalls on
// long_scoreboard, profiles/r1_ncu_iw.md).  Out-of-image halo cells are zero-filled by the TMA unit.
// With FUSE3 the new direction p = z + beta*pOld is formed once per staged cell in shared memory.
// Each thread marches down a 4-pixel column strip: the vertical pair term is evaluated once and
// credited to both of its pixels (the J^T J pair block is antisymmetric in the offset rows), the
// horizontal pairs are evaluated one-sided from both ends so no cross-lane exchange is needed.
// Shape: 64 x IW_TH pixel tiles, IW_STAGES tiles in flight per CTA, IW_STRIP rows per thread (256 threads = 64 columns x IW_TH / IW_STRIP
// strips).  Measured in the bench loop at 2048^2 (fused kernel, tools/build_variants.sh): 16/2/4 (2 CTAs per SM) 59.9 us, 8/2/2 61.1,
// 8/3/2 59.3, 8/4/2 59.5, 8/6/2 69.8, 32/2/8 58.6, 16/4/4 59.6, 16/3/4 (one CTA per SM, two tiles ahead) 55.4-56.1 -- the default.
#ifndef IW_TH
#define IW_TH 16
#define IW_STAGES 3
#define IW_STRIP 4
#endif
constexpr int kTW = 64, kTH = IW_TH, kStagesSingle = IW_STAGES, kStagesMulti = 2, kStrip = IW_STRIP;
// Pipeline depth per use: one GPU = IW_STAGES (3: one CTA per SM, two tiles ahead).  Multi-rank slabs = 2 stages, two CTAs per SM: there the
// last CTA first pushes the halo rows to the neighbours (NVLink stores + a system-scope release, several microseconds); with one CTA per
// SM that SM idles meanwhile and its CTA becomes the kernel's straggler (measured at 4 ranks: 50.3 instead of 45.2 us per PCG step).
// TMA rule (measured, tools/tma_probe.cu): the innermost box coordinate must start on a 16-byte
// boundary, so 2-channel rows are staged from pixel x0-2 and 1-channel rows from pixel x0-4.
constexpr int kRowO = 2 * (kTW + 4);      // floats per staged row of a 2-channel image (pixels x0-2 .. x0+TW+1)
constexpr int kRowA = kTW + 8;            // floats per staged row of a 1-channel image (pixels x0-4 .. x0+TW+3)
constexpr int kRows = kTH + 2;
constexpr int kBytesO = kRows * kRowO * 4, kBytesA = kRows * kRowA * 4, kBytesG = kRows * (kTW + 2) * 16, kBytesF = kTH * kTW;
constexpr int align128(int v) { return (v + 127) / 128 * 128; }
constexpr int kOffPO = 0, kOffPA = kOffPO + align128(kBytesO), kOffG = kOffPA + align128(kBytesA), kOffF = kOffG + align128(kBytesG);
constexpr int kOffZO = kOffF + align128(kBytesF), kOffZA = kOffZO + align128(kBytesO);
constexpr int kStageBytesPlain = kOffZO, kStageBytesFused = kOffZA + align128(kBytesA);
static_assert(kBlock == kTW * (kTH / kStrip), "thread mapping: one thread per column strip");

struct IWMaps { CUtensorMap pO, pA, zO, zA, geom, flags; };

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count)); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoad2D(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smemAddr(dst)), "l"(map), "r"(smemAddr(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmaLoad3D(void* dst, const CUtensorMap* map, uint64_t* bar, int c, int x, int y) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smemAddr(dst)), "l"(map), "r"(smemAddr(bar)), "r"(c), "r"(x), "r"(y) : "memory");
}

struct IWPoint { float c, s, ux, uy, px, py, pa; };   // one staged cell: geom + direction
__device__ __forceinline__ IWPoint iwLoadPoint(const float4* sG, const float* sPO, const float* sPA, int r, int lx) {
    const float4 g = sG[r * (kTW + 2) + lx + 1];
    const float2 p = *reinterpret_cast<const float2*>(&sPO[r * kRowO + 2 * (lx + 2)]);
    const float a = sPA[r * kRowA + lx + 4];
    return IWPoint{ g.x, g.y, g.z, g.w, p.x, p.y, a };
}
// pair term for the pair (i, j):  t is credited +t to Ap_O(i) and -t to Ap_O(j); Ai / Aj go to Ap_a(i) / Ap_a(j)
__device__ __forceinline__ void iwPairBoth(const IWPoint& i, const IWPoint& j, float& tx, float& ty, float& Ai, float& Aj) {
    const float dx = i.ux - j.ux, dy = i.uy - j.uy;
    const float eix = -i.s * dx - i.c * dy, eiy = i.c * dx - i.s * dy;
    const float ejx = -j.s * dx - j.c * dy, ejy = j.c * dx - j.s * dy;
    const float px = i.px - j.px, py = i.py - j.py;
    tx = 2.f * px - eix * i.pa - ejx * j.pa; ty = 2.f * py - eiy * i.pa - ejy * j.pa;
    Ai = (eix * eix + eiy * eiy) * i.pa - (eix * px + eiy * py);
    Aj = (ejx * ejx + ejy * ejy) * j.pa - (ejx * px + ejy * py);
}
__device__ __forceinline__ void iwPairOne(const IWPoint& i, const IWPoint& j, float& tx, float& ty, float& Ai) {
    const float dx = i.ux - j.ux, dy = i.uy - j.uy;
    const float eix = -i.s * dx - i.c * dy, eiy = i.c * dx - i.s * dy;
    const float ejx = -j.s * dx - j.c * dy, ejy = j.c * dx - j.s * dy;
    const float px = i.px - j.px, py = i.py - j.py;
    tx = 2.f * px - eix * i.pa - ejx * j.pa; ty = 2.f * py - eiy * i.pa - ejy * j.pa;
    Ai = (eix * eix + eiy * eiy) * i.pa - (eix * px + eiy * py);
}

template <bool FUSE3, bool LM, int kStages>
__global__ void __launch_bounds__(kBlock) k_iwApplyJTJTma(IWData<float> D, const __grid_constant__ IWMaps M, ReduceWorkspace ws, int dotSlot, int rzNew,
                                                          int rzOld, float* __restrict__ pOut, float* __restrict__ out, const float* __restrict__ CtC,
                                                          int tilesX, int numTiles, const float* __restrict__ haloSrc, const HaloPlan* __restrict__ hpp) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[kStages];
    constexpr int stageBytes = FUSE3 ? kStageBytesFused : kStageBytesPlain;
    const int tid = threadIdx.x;
    pdlTrigger();
    if (tid == 0) {   // prologue that does not depend on the previous kernel overlaps its tail (PDL)
        for (int s = 0; s < kStages; ++s) mbarInit(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdlWait();
    __syncthreads();
    // Multi-GPU (hp.nSeg > 0): this kernel is also the halo exchange.  The last CTA first stores this rank's
    // boundary rows of haloSrc (z when FUSE3, else p) into the neighbours' halo rows over NVLink and releases
    // their flags; every CTA waits for the incoming flag only before staging a tile that touches a halo row,
    // and the tile order is rotated so those tiles come mid-kernel, when the flags have long arrived.
    const bool halo = hpp != nullptr;
    const unsigned long long haloE = halo ? haloEpochNow(ws) : 0ull;
    unsigned long long* myFlagUp = nullptr; unsigned long long* myFlagDn = nullptr;   // non-null: a neighbour exists on that side
    if (halo) { if (hpp->peerVec[0]) myFlagUp = hpp->myFlag[0]; if (hpp->peerVec[1]) myFlagDn = hpp->myFlag[1]; }
    const int tileRot = halo ? ((numTiles / tilesX) / 2) * tilesX : 0;
    auto issue = [&](int seq, int stage) {   // one thread
        int tile = seq + tileRot; if (tile >= numTiles) tile -= numTiles;
        unsigned char* base = smem + (size_t)stage * stageBytes;
        const int x0 = (tile % tilesX) * kTW, y0 = (tile / tilesX) * kTH;
        if (halo) {
            bool waited = false;
            if (myFlagUp && y0 - 1 <= D.y0 - 1) { SpinGuard g; while (ldAcquireSys(myFlagUp) < haloE) g.tick(); waited = true; }
            if (myFlagDn && y0 + kTH >= D.y1) { SpinGuard g; while (ldAcquireSys(myFlagDn) < haloE) g.tick(); waited = true; }
            if (waited) asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        mbarExpectTx(&bars[stage], (unsigned)(kBytesO + kBytesA + kBytesG + kBytesF + (FUSE3 ? kBytesO + kBytesA : 0)));
        tmaLoad2D(base + kOffPO, &M.pO, &bars[stage], 2 * (x0 - 2), y0 - 1);
        tmaLoad2D(base + kOffPA, &M.pA, &bars[stage], x0 - 4, y0 - 1);
        tmaLoad3D(base + kOffG, &M.geom, &bars[stage], 0, x0 - 1, y0 - 1);
        tmaLoad2D(base + kOffF, &M.flags, &bars[stage], x0, y0);
        if (FUSE3) {
            tmaLoad2D(base + kOffZO, &M.zO, &bars[stage], 2 * (x0 - 2), y0 - 1);
            tmaLoad2D(base + kOffZA, &M.zA, &bars[stage], x0 - 4, y0 - 1);
        }
    };
    if (halo && blockIdx.x == gridDim.x - 1) {
        const HaloPlan& hp = *hpp;
        for (int s = 0; s < hp.nSeg; ++s) {
            float* dst = reinterpret_cast<float*>(hp.peerVec[hp.dir[s]]);
            if (!dst) continue;
            const float4* src4 = reinterpret_cast<const float4*>(haloSrc + hp.srcOff[s]); float4* dst4 = reinterpret_cast<float4*>(dst + hp.dstOff[s]);
            for (long i = tid; i < hp.len[s] / 4; i += kBlock) dst4[i] = src4[i];
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) for (int d = 0; d < 2; ++d) if (hp.peerVec[d]) stReleaseSys(hp.peerFlag[d], haloE);
    }
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { const int t = blockIdx.x + s * gridDim.x; if (t < numTiles) issue(t, s); }
    }
    // beta is awaited only now, with the first tiles already in flight: nothing staged above depends on it, and on several GPUs it
    // arrives through the NVLink mailboxes a few microseconds after the kernel starts
    float beta = 0.f;
    if (FUSE3) {
        double drn, dro; loadScalars2(ws, rzNew, rzOld, drn, dro);
        const float rn = (float)drn, ro = (float)dro;
        if (ro > 0.f) beta = rn / ro;
    }
    const float w2 = D.w * D.w, wf2 = D.wf * D.wf;
    const int lx = tid & (kTW - 1), ly0 = (tid >> 6) * kStrip;   // 256 threads = 64 columns x 4 strips of 4 rows
    float acc = 0.f;
    int it = 0;
    for (int seq = blockIdx.x; seq < numTiles; seq += gridDim.x, ++it) {
        int tile = seq + tileRot; if (tile >= numTiles) tile -= numTiles;
        const int stage = it % kStages;
        const unsigned parity = (unsigned)((it / kStages) & 1);
        unsigned char* base = smem + (size_t)stage * stageBytes;
        const int x0 = (tile % tilesX) * kTW, y0 = (tile / tilesX) * kTH;
        mbarWait(&bars[stage], parity);
        float* sPO = reinterpret_cast<float*>(base + kOffPO);
        float* sPA = reinterpret_cast<float*>(base + kOffPA);
        const float4* sG = reinterpret_cast<const float4*>(base + kOffG);
        const unsigned char* sF = base + kOffF;
        if (FUSE3) {   // p = z + beta * pOld for every staged cell (halo included), in place
            const float4* zO4 = reinterpret_cast<const float4*>(base + kOffZO); float4* pO4 = reinterpret_cast<float4*>(sPO);
#pragma unroll 2
            for (int i = tid; i < kRows * kRowO / 4; i += kBlock) { const float4 zv = zO4[i]; float4 pv = pO4[i]; pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y; pv.z = zv.z + beta * pv.z; pv.w = zv.w + beta * pv.w; pO4[i] = pv; }
            const float4* zA4 = reinterpret_cast<const float4*>(base + kOffZA); float4* pA4 = reinterpret_cast<float4*>(sPA);
            for (int i = tid; i < kRows * kRowA / 4; i += kBlock) { const float4 zv = zA4[i]; float4 pv = pA4[i]; pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y; pv.z = zv.z + beta * pv.z; pv.w = zv.w + beta * pv.w; pA4[i] = pv; }
            __syncthreads();
        }
        const int gx = x0 + lx;
        // staged row r holds image row y0 - 1 + r; pixel (lx, ly) is the cell (r = ly + 1)
        IWPoint up = iwLoadPoint(sG, sPO, sPA, ly0, lx);
        IWPoint cur = iwLoadPoint(sG, sPO, sPA, ly0 + 1, lx);
        float cx = 0.f, cy = 0.f, ca = 0.f;   // contribution of the pair with the pixel above, carried down the strip
        {
            const unsigned f0 = sF[ly0 * kTW + lx];
            float tx, ty, Ai, Aj; iwPairBoth(up, cur, tx, ty, Ai, Aj);
            const float v = (f0 & 8u) ? 1.f : 0.f;   // bit 3: neighbour (0,-1)
            cx = -v * tx; cy = -v * ty; ca = v * Aj;
        }
#pragma unroll
        for (int q = 0; q < kStrip; ++q) {
            const int ly = ly0 + q, r = ly + 1, gy = y0 + ly;
            const unsigned fl = sF[ly * kTW + lx];
            const IWPoint dn = iwLoadPoint(sG, sPO, sPA, r + 1, lx);
            const IWPoint lf = iwLoadPoint(sG, sPO, sPA, r, lx - 1);
            const IWPoint rt = iwLoadPoint(sG, sPO, sPA, r, lx + 1);
            float ax = cx, ay = cy, aa = ca;
            float tx, ty, Ai, Aj;
            iwPairBoth(cur, dn, tx, ty, Ai, Aj);
            const float vd = (fl & 4u) ? 1.f : 0.f;     // bit 2: neighbour (0,+1)
            ax += vd * tx; ay += vd * ty; aa += vd * Ai;
            cx = -vd * tx; cy = -vd * ty; ca = vd * Aj;
            iwPairOne(cur, rt, tx, ty, Ai);
            const float vr = (fl & 1u) ? 1.f : 0.f;     // bit 0: neighbour (+1,0)
            ax += vr * tx; ay += vr * ty; aa += vr * Ai;
            iwPairOne(cur, lf, tx, ty, Ai);
            const float vl = (fl & 2u) ? 1.f : 0.f;     // bit 1: neighbour (-1,0)
            ax += vl * tx; ay += vl * ty; aa += vl * Ai;
            const float vf = (fl & 16u) ? wf2 : 0.f;
            float2 oO = make_float2(w2 * ax + vf * cur.px, w2 * ay + vf * cur.py);
            float oA = w2 * aa;
            if (gy < D.y0 || gy >= D.y1) { oO = make_float2(0.f, 0.f); oA = 0.f; }   // halo row: direction only, no Ap
            if (gx < D.W && gy < D.H) {
                const long i = (long)gy * D.W + gx;
                if (LM) {
                    const float2 cO = reinterpret_cast<const float2*>(CtC)[i];
                    oO.x += cO.x * cur.px; oO.y += cO.y * cur.py; oA += CtC[D.offA + i] * cur.pa;
                }
                acc += cur.px * oO.x + cur.py * oO.y + cur.pa * oA;
                reinterpret_cast<float2*>(out)[i] = oO; out[D.offA + i] = oA;
                if (FUSE3) { reinterpret_cast<float2*>(pOut)[i] = make_float2(cur.px, cur.py); pOut[D.offA + i] = cur.pa; }
            }
            cur = dn;
        }
        __syncthreads();   // every thread is done with this stage
        if (tid == 0) {
            const int next = seq + kStages * gridDim.x;
            if (next < numTiles) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(next, stage); }
        }
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}

// ------------------------------------------------------------------------------------ whole PCG loop, persistent
// All L iterations of PCGStep1/2/3 (GN, solverGPUGaussNewton.t:421-550) in ONE cooperative launch.  For images whose
// PCG vectors fit in L2 the kernel-per-step loop is bound by launch latency and kernel ramp/tail (15 us per step for
// 512x512 with two kernels per step), not by bandwidth; here every CTA keeps a fixed set of 32x8 tiles and the
// iterations are separated by two grid-wide all-reduces, which are needed anyway (p.Ap before alpha, z.r before beta):
//   phase A   p = z + beta*pOld for the tile and its 1-pixel ring (staged in shared memory; iteration 0 takes p as
//             is), Ap = J^T J p, partial p.Ap                      -> all-reduce #1 -> alpha
//   phase B   delta += alpha p, r -= alpha Ap, z = pre*r, partial z.r   -> all-reduce #2 -> beta
// p is double-buffered like in the fused kernel-per-step path (neighbours in other CTAs read pOld while the new
// centre value is written to the other buffer).  Everything other CTAs rewrite during the kernel is read with
// ld.global.cg (L2): L1 is not coherent across SMs.  The arithmetic per pixel is that of k_iwApplyJTJ / k_pcgStep2 /
// k_pcgStep3, so the results agree with the kernel-per-step loop up to the summation order of the dot products.
__device__ __forceinline__ float ldcgT(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcgT(const double* p) { return __ldcg(p); }
__device__ __forceinline__ Vec2<float> ldcgV(const Vec2<float>* p) { const float2 v = __ldcg(reinterpret_cast<const float2*>(p)); return Vec2<float>{ v.x, v.y }; }
__device__ __forceinline__ Vec2<double> ldcgV(const Vec2<double>* p) { const double2 v = __ldcg(reinterpret_cast<const double2*>(p)); return Vec2<double>{ v.x, v.y }; }
__device__ __forceinline__ void stcgT(float* p, float v) { __stcg(p, v); }
__device__ __forceinline__ void stcgT(double* p, double v) { __stcg(p, v); }
__device__ __forceinline__ void stcgV(Vec2<float>* p, Vec2<float> v) { __stcg(reinterpret_cast<float2*>(p), make_float2(v.x, v.y)); }
__device__ __forceinline__ void stcgV(Vec2<double>* p, Vec2<double> v) { __stcg(reinterpret_cast<double2*>(p), make_double2(v.x, v.y)); }

constexpr int kPW = kTileW + 2, kPH = kTileH + 2;   // staged cells per tile (1-pixel ring)

template <class T>
__global__ void __launch_bounds__(kBlock) k_iwPcgLoop(IWData<T> D, ReduceWorkspace ws, LoopSync* S, unsigned long long k0, int L,
                                                      T* delta, T* r, T* z, T* P0, T* P1, T* Ap, const T* __restrict__ pre) {
    __shared__ T sPx[kPH][kPW], sPy[kPH][kPW], sPa[kPH][kPW];
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    double* partials = ws.partials + (size_t)kPAp * kMaxPartials;   // the two reductions alternate, one buffer is enough
    unsigned long long k = k0;
    double dRz = ws.scalars[kRz0], dRzPrev = 0.0, dPAp = 0.0;        // r.z of PCGInit1 (previous kernel on the stream)
    T* cur = P0; T* alt = P1;
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    for (int l = 0; l < L; ++l) {
        T beta = T(0);
        if (l > 0) { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }   // PCGStep3 of iteration l-1 (:541-546)
        // ---------------- phase A
        T acc = T(0);
        for (int tile = blockIdx.x; tile < D.numTiles; tile += gridDim.x) {
            const int x0 = (tile % D.tilesX) * kTileW, y0 = (tile / D.tilesX) * kTileH;
            __syncthreads();   // previous tile's readers are done with the staging buffers
            for (int c = tid; c < kPW * kPH; c += kBlock) {
                const int cy = c / kPW, cx = c - cy * kPW;
                const int gx = x0 - 1 + cx, gy = y0 - 1 + cy;
                T px = T(0), py = T(0), pa = T(0);
                if (gx >= 0 && gx < D.W && gy >= 0 && gy < D.H) {
                    const long j = (long)gy * D.W + gx;
                    const Vec2<T> po = ldcgV(reinterpret_cast<const Vec2<T>*>(cur) + j);
                    pa = ldcgT(cur + D.offA + j);
                    px = po.x; py = po.y;
                    if (l > 0) {
                        const Vec2<T> zo = ldcgV(reinterpret_cast<const Vec2<T>*>(z) + j);
                        const T za = ldcgT(z + D.offA + j);
                        px = zo.x + beta * px; py = zo.y + beta * py; pa = za + beta * pa;
                    }
                }
                sPx[cy][cx] = px; sPy[cy][cx] = py; sPa[cy][cx] = pa;
            }
            __syncthreads();
            const int x = x0 + lx, y = y0 + ly;
            if (x < D.W && y < D.H) {
                const long i = (long)y * D.W + x;
                const unsigned fl = D.flags[i];
                const T pix = sPx[ly + 1][lx + 1], piy = sPy[ly + 1][lx + 1], pai = sPa[ly + 1][lx + 1];
                if (l > 0) { stcgV(reinterpret_cast<Vec2<T>*>(alt) + i, Vec2<T>{ pix, piy }); stcgT(alt + D.offA + i, pai); }
                T ox = T(0), oy = T(0), oa = T(0);
                if (fl & 32u) {
                    const Vec4<T> gi = D.geom[i];
                    T ax = T(0), ay = T(0), aa = T(0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (fl & (1u << q)) {
                            const int ddx = q == 0 ? 1 : (q == 1 ? -1 : 0), ddy = q == 2 ? 1 : (q == 3 ? -1 : 0);
                            const Vec4<T> gj = D.geom[i + (long)ddy * D.W + ddx];
                            const T pjx = sPx[ly + 1 + ddy][lx + 1 + ddx], pjy = sPy[ly + 1 + ddy][lx + 1 + ddx], paj = sPa[ly + 1 + ddy][lx + 1 + ddx];
                            const T dx = gi.z - gj.z, dy = gi.w - gj.w;
                            const T eix = -gi.y * dx - gi.x * dy, eiy = gi.x * dx - gi.y * dy;
                            const T ejx = -gj.y * dx - gj.x * dy, ejy = gj.x * dx - gj.y * dy;
                            const T px = pix - pjx, py = piy - pjy;
                            ax += T(2) * px - eix * pai - ejx * paj;
                            ay += T(2) * py - eiy * pai - ejy * paj;
                            aa += (eix * eix + eiy * eiy) * pai - (eix * px + eiy * py);
                        }
                    }
                    ox = w2 * ax; oy = w2 * ay; oa = w2 * aa;
                    if (fl & 16u) { ox += wf2 * pix; oy += wf2 * piy; }
                    acc += pix * ox + piy * oy + pai * oa;
                }
                stcgV(reinterpret_cast<Vec2<T>*>(Ap) + i, Vec2<T>{ ox, oy }); stcgT(Ap + D.offA + i, oa);
            }
        }
        if (l > 0) { T* t = cur; cur = alt; alt = t; }
        dPAp = gridAllReduce<kBlock>((double)acc, partials, S, ++k);
        // ---------------- phase B (PCGStep2 :446-489)
        T alpha = T(0);
        { const T num = (T)dRz, den = (T)dPAp; if (den > T(0)) alpha = num / den; }
        acc = T(0);
        for (int tile = blockIdx.x; tile < D.numTiles; tile += gridDim.x) {
            const int x = (tile % D.tilesX) * kTileW + lx, y = (tile / D.tilesX) * kTileH + ly;
            if (x < D.W && y < D.H) {
                const long i = (long)y * D.W + x;
                Vec2<T> dO = ldcgV(reinterpret_cast<const Vec2<T>*>(delta) + i); T dA = ldcgT(delta + D.offA + i);
                const Vec2<T> pO = ldcgV(reinterpret_cast<const Vec2<T>*>(cur) + i); const T pA = ldcgT(cur + D.offA + i);
                Vec2<T> rO = ldcgV(reinterpret_cast<const Vec2<T>*>(r) + i); T rA = ldcgT(r + D.offA + i);
                const Vec2<T> aO = ldcgV(reinterpret_cast<const Vec2<T>*>(Ap) + i); const T aA = ldcgT(Ap + D.offA + i);
                const Vec2<T> qO = reinterpret_cast<const Vec2<T>*>(pre)[i]; const T qA = pre[D.offA + i];
                dO.x = dO.x + alpha * pO.x; dO.y = dO.y + alpha * pO.y; dA = dA + alpha * pA;
                rO.x = rO.x - alpha * aO.x; rO.y = rO.y - alpha * aO.y; rA = rA - alpha * aA;
                const Vec2<T> zO{ qO.x * rO.x, qO.y * rO.y }; const T zA = qA * rA;
                acc += zO.x * rO.x; acc += zO.y * rO.y; acc += zA * rA;
                stcgV(reinterpret_cast<Vec2<T>*>(delta) + i, dO); stcgT(delta + D.offA + i, dA);
                stcgV(reinterpret_cast<Vec2<T>*>(r) + i, rO); stcgT(r + D.offA + i, rA);
                stcgV(reinterpret_cast<Vec2<T>*>(z) + i, zO); stcgT(z + D.offA + i, zA);
            }
        }
        dRzPrev = dRz;
        dRz = gridAllReduce<kBlock>((double)acc, partials, S, ++k);
    }
    // PCGStep3 of the last iteration, in place (nobody reads neighbours any more)
    if (L > 0) {
        T beta = T(0);
        { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
        for (int tile = blockIdx.x; tile < D.numTiles; tile += gridDim.x) {
            const int x = (tile % D.tilesX) * kTileW + lx, y = (tile / D.tilesX) * kTileH + ly;
            if (x < D.W && y < D.H) {
                const long i = (long)y * D.W + x;
                const Vec2<T> zO = ldcgV(reinterpret_cast<const Vec2<T>*>(z) + i); const T zA = ldcgT(z + D.offA + i);
                const Vec2<T> pO = ldcgV(reinterpret_cast<const Vec2<T>*>(cur) + i); const T pA = ldcgT(cur + D.offA + i);
                stcgV(reinterpret_cast<Vec2<T>*>(cur) + i, Vec2<T>{ zO.x + beta * pO.x, zO.y + beta * pO.y }); stcgT(cur + D.offA + i, zA + beta * pA);
            }
        }
        if (blockIdx.x == 0 && tid == 0) {   // leave the scalars as the kernel-per-step loop does (iteration l reads rz[l&1], writes rz[(l+1)&1])
            ws.scalars[kPAp] = dPAp; ws.scalars[L & 1] = dRz; ws.scalars[(L + 1) & 1] = dRzPrev;
        }
    }
}

// ------------------------------------------------------------------------------------ whole PCG loop, state on chip
// Same algorithm as k_iwPcgLoop, but the loop state does not leave the SM: the image is cut into slabs of R full rows,
// one slab per CTA (1024 threads, one CTA per SM, cooperative launch), and p, r, pre, the packed geometry and the flag
// bytes of the slab live in shared memory for all L iterations; z, Ap and delta of a thread's (at most KMAX) pixels live
// in registers.  Per iteration only three things touch L2: the two boundary rows of z each slab publishes for its
// neighbours (phase B), the two ring rows of z it reads back (phase A: p_ring = z_ring + beta * p_ring, the ring's old p
// is already on chip), and the two all-reduces.  Applies when ceil(H / #SMs) rows fit the shared memory of one SM.
template <class T, int KMAX>
__global__ void __launch_bounds__(1024, 1) k_iwPcgLoopResident(IWData<T> D, ReduceWorkspace ws, unsigned long long* llWords, unsigned long long* arrivals, unsigned long long k0, int L, int R,
                                                               T* delta, T* r, T* z, T* p, T* Ap, const T* __restrict__ pre) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int NT = 1024;
    const int W = D.W, H = D.H, tid = threadIdx.x;
    const int y0 = blockIdx.x * R;
    const int rows = min(R, H - y0);                   // >= 1: the host launches ceil(H / R) CTAs
    const int nOwn = rows * W, nRing = (R + 2) * W, nCap = R * W;
    // planes with a ring row above and below (local row lr <-> image row y0 - 1 + lr): direction and packed geometry
    T* sPx = reinterpret_cast<T*>(smemRaw); T* sPy = sPx + nRing; T* sPa = sPy + nRing;
    T* sGc = sPa + nRing; T* sGs = sGc + nRing; T* sGx = sGs + nRing; T* sGy = sGx + nRing;
    // planes of the owned rows only: residual, preconditioner (Offset part is isotropic), flag bytes
    T* sRx = sGy + nRing; T* sRy = sRx + nCap; T* sRa = sRy + nCap; T* sQo = sRa + nCap; T* sQa = sQo + nCap;
    unsigned char* sF = reinterpret_cast<unsigned char*>(sQa + nCap);
    const Vec2<T>* pO = reinterpret_cast<const Vec2<T>*>(p);
    for (int c = tid; c < (rows + 2) * W; c += NT) {
        const int lr = c / W, x = c - lr * W, gy = y0 - 1 + lr;
        T px = T(0), py = T(0), pa = T(0); Vec4<T> g{ T(0), T(0), T(0), T(0) };
        if (gy >= 0 && gy < H) { const long j = (long)gy * W + x; const Vec2<T> po = pO[j]; px = po.x; py = po.y; pa = p[D.offA + j]; g = D.geom[j]; }
        sPx[c] = px; sPy[c] = py; sPa[c] = pa; sGc[c] = g.x; sGs[c] = g.y; sGx[c] = g.z; sGy[c] = g.w;
    }
    T zx[KMAX], zy[KMAX], za[KMAX], ax[KMAX], ay[KMAX], aa[KMAX], dx[KMAX], dy[KMAX], da[KMAX];
    bool edgeRow[KMAX];                                // pixel lies on the slab's first or last row: its z is read by a neighbour
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int c = tid + q * NT;
        zx[q] = zy[q] = za[q] = ax[q] = ay[q] = aa[q] = dx[q] = dy[q] = da[q] = T(0); edgeRow[q] = false;
        if (c < nOwn) {
            const long i = (long)y0 * W + c;
            const Vec2<T> ro = reinterpret_cast<const Vec2<T>*>(r)[i]; sRx[c] = ro.x; sRy[c] = ro.y; sRa[c] = r[D.offA + i];
            sQo[c] = reinterpret_cast<const Vec2<T>*>(pre)[i].x; sQa[c] = pre[D.offA + i];
            sF[c] = D.flags[i];
            const Vec2<T> dO = reinterpret_cast<const Vec2<T>*>(delta)[i]; dx[q] = dO.x; dy[q] = dO.y; da[q] = delta[D.offA + i];
            const int lr = c / W; edgeRow[q] = (lr == 0 || lr == rows - 1);
        }
    }
    __syncthreads();
    unsigned long long k = k0;
    double dRz = ws.scalars[kRz0], dRzPrev = 0.0, dPAp = 0.0;
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    for (int l = 0; l < L; ++l) {
        // ---------------- phase A: p = z + beta p (own cells and ring rows), Ap = J^T J p
        if (l > 0) {
            T beta = T(0);
            { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
#pragma unroll
            for (int q = 0; q < KMAX; ++q) {
                const int c = tid + q * NT;
                if (c < nOwn) { const int cc = c + W; sPx[cc] = zx[q] + beta * sPx[cc]; sPy[cc] = zy[q] + beta * sPy[cc]; sPa[cc] = za[q] + beta * sPa[cc]; }
            }
            for (int c = tid; c < 2 * W; c += NT) {
                const int side = c >= W ? 1 : 0, x = c - side * W, gy = side ? y0 + rows : y0 - 1;
                if (gy >= 0 && gy < H) {
                    const long j = (long)gy * W + x;
                    const Vec2<T> zo = ldcgV(reinterpret_cast<const Vec2<T>*>(z) + j); const T zaa = ldcgT(z + D.offA + j);
                    const int cc = side ? (rows + 1) * W + x : x;
                    sPx[cc] = zo.x + beta * sPx[cc]; sPy[cc] = zo.y + beta * sPy[cc]; sPa[cc] = zaa + beta * sPa[cc];
                }
            }
            __syncthreads();
        }
        T acc = T(0);
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            const int c = tid + q * NT;
            T ox = T(0), oy = T(0), oa = T(0);
            if (c < nOwn) {
                const int cc = c + W;
                const unsigned fl = sF[c];
                if (fl & 32u) {
                    const T pix = sPx[cc], piy = sPy[cc], pai = sPa[cc];
                    const T gc = sGc[cc], gs = sGs[cc], gux = sGx[cc], guy = sGy[cc];
                    T sx = T(0), sy = T(0), sa = T(0);
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        if (fl & (1u << nb)) {
                            const int jj = cc + (nb == 0 ? 1 : (nb == 1 ? -1 : (nb == 2 ? W : -W)));
                            const T jc = sGc[jj], js = sGs[jj];
                            const T ddx = gux - sGx[jj], ddy = guy - sGy[jj];
                            const T eix = -gs * ddx - gc * ddy, eiy = gc * ddx - gs * ddy;
                            const T ejx = -js * ddx - jc * ddy, ejy = jc * ddx - js * ddy;
                            const T paj = sPa[jj];
                            const T px = pix - sPx[jj], py = piy - sPy[jj];
                            sx += T(2) * px - eix * pai - ejx * paj;
                            sy += T(2) * py - eiy * pai - ejy * paj;
                            sa += (eix * eix + eiy * eiy) * pai - (eix * px + eiy * py);
                        }
                    }
                    ox = w2 * sx; oy = w2 * sy; oa = w2 * sa;
                    if (fl & 16u) { ox += wf2 * pix; oy += wf2 * piy; }
                    acc += pix * ox + piy * oy + pai * oa;
                }
            }
            ax[q] = ox; ay[q] = oy; aa[q] = oa;
        }
        dPAp = gridAllReduceLL<NT>((double)acc, llWords, arrivals, ++k, false);   // nothing but the scalar crosses CTAs here
        // ---------------- phase B (PCGStep2)
        T alpha = T(0);
        { const T num = (T)dRz, den = (T)dPAp; if (den > T(0)) alpha = num / den; }
        acc = T(0);
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            const int c = tid + q * NT;
            if (c < nOwn) {
                const int cc = c + W;
                dx[q] = dx[q] + alpha * sPx[cc]; dy[q] = dy[q] + alpha * sPy[cc]; da[q] = da[q] + alpha * sPa[cc];
                const T rx = sRx[c] - alpha * ax[q], ry = sRy[c] - alpha * ay[q], ra = sRa[c] - alpha * aa[q];
                sRx[c] = rx; sRy[c] = ry; sRa[c] = ra;
                const T qo = sQo[c];
                zx[q] = qo * rx; zy[q] = qo * ry; za[q] = sQa[c] * ra;
                acc += zx[q] * rx; acc += zy[q] * ry; acc += za[q] * ra;
                if (edgeRow[q]) {
                    const long i = (long)y0 * W + c;
                    stcgV(reinterpret_cast<Vec2<T>*>(z) + i, Vec2<T>{ zx[q], zy[q] }); stcgT(z + D.offA + i, za[q]);
                }
            }
        }
        dRzPrev = dRz;
        dRz = gridAllReduceLL<NT>((double)acc, llWords, arrivals, ++k, true);      // releases the edge rows of z to the neighbouring slabs
    }
    // PCGStep3 of the last iteration, then the state goes back to the plan's vectors
    T beta = T(0);
    if (L > 0) { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int c = tid + q * NT;
        if (c < nOwn) {
            const int cc = c + W; const long i = (long)y0 * W + c;
            T px = sPx[cc], py = sPy[cc], pa = sPa[cc];
            if (L > 0) { px = zx[q] + beta * px; py = zy[q] + beta * py; pa = za[q] + beta * pa; }
            reinterpret_cast<Vec2<T>*>(p)[i] = Vec2<T>{ px, py }; p[D.offA + i] = pa;
            reinterpret_cast<Vec2<T>*>(delta)[i] = Vec2<T>{ dx[q], dy[q] }; delta[D.offA + i] = da[q];
            reinterpret_cast<Vec2<T>*>(r)[i] = Vec2<T>{ sRx[c], sRy[c] }; r[D.offA + i] = sRa[c];
            reinterpret_cast<Vec2<T>*>(z)[i] = Vec2<T>{ zx[q], zy[q] }; z[D.offA + i] = za[q];
            reinterpret_cast<Vec2<T>*>(Ap)[i] = Vec2<T>{ ax[q], ay[q] }; Ap[D.offA + i] = aa[q];
        }
    }
    if (L > 0 && blockIdx.x == 0 && tid == 0) { ws.scalars[kPAp] = dPAp; ws.scalars[L & 1] = dRz; ws.scalars[(L + 1) & 1] = dRzPrev; }
}

// ------------------------------------------------------------------------------------ whole PCG loop, state on chip, 2-D tiles
// k_iwPcgLoopResident with tiles of tw x th pixels instead of slabs of full rows: the ring is the tile's perimeter (all
// four sides), so wide images whose row slabs would be all ring still fit (the shape the
// multi-GPU version needs).  OPT-IN (OPTB_RESIDENT_TILE=WxH) and NOT yet run on hardware: written after the round's GPU
// budget was spent; the validated default is the row-slab kernel above.
// MULTI (one process per GPU, row-slab decomposition, every rank launches this kernel collectively): the tiles cover the
// rank's OWNED rows [D.y0, D.y1); ring cells in a halo row read the z the neighbouring rank stored there; tiles on the first /
// last owned row also store their edge row of z into the neighbour's halo row over NVLink; the reductions span all ranks
// (gridAllReduceLLMulti); and because the per-step path keeps X consistent on halo rows by recomputing delta there
// (DESIGN.md section 6), boundary tiles accumulate delta for their ring row too ("shadow delta", sH*).
// Rank boundary of the collective loop kernel: z of my first / last owned row travels to the neighbouring rank as LL words (llStore /
// llLoad, common.cuh) in that rank's peer scratch: 3 scalars per pixel (Offset.x, Offset.y, Angle) at word index (3 x + c) * LLWords<T>::N.
// toUp / toDn: where I write (the upper neighbour's "from below" row / the lower neighbour's "from above" row; nullptr: no neighbour);
// fromUp / fromDn: my own two rows, written by the neighbours.
template <class T> struct IWTileMulti { unsigned long long* toUp; unsigned long long* toDn; const unsigned long long* fromUp; const unsigned long long* fromDn; };
constexpr int kLoopMailboxSlot = 8;   // spare scalar slot of CommBlock::words used by the persistent loop's own sequence numbers
template <class T, int KMAX, bool MULTI>
__global__ void __launch_bounds__(1024, 1) k_iwPcgLoopTile(IWData<T> D, ReduceWorkspace ws, unsigned long long* llWords, unsigned long long* arrivals,
                                                           unsigned long long k0, int L, int tw, int th, int tilesAcross, IWTileMulti<T> M,
                                                           T* delta, T* r, T* z, T* p, T* Ap, const T* __restrict__ pre) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int NT = 1024;
    const int W = D.W, H = D.H, tid = threadIdx.x;
    const int tx0 = (blockIdx.x % tilesAcross) * tw, ty0 = D.y0 + (blockIdx.x / tilesAcross) * th;
    const int twE = min(tw, W - tx0), thE = min(th, D.y1 - ty0);   // clipped tile (>= 1 x 1 by construction); single GPU: [D.y0, D.y1) = [0, H)
    const int pw = tw + 2;                                        // pitch of the ring-included planes
    const int nRing = (th + 2) * pw, nCap = tw * th;
    T* sPx = reinterpret_cast<T*>(smemRaw); T* sPy = sPx + nRing; T* sPa = sPy + nRing;
    T* sGc = sPa + nRing; T* sGs = sGc + nRing; T* sGx = sGs + nRing; T* sGy = sGx + nRing;
    T* sRx = sGy + nRing; T* sRy = sRx + nCap; T* sRa = sRy + nCap; T* sQo = sRa + nCap; T* sQa = sQo + nCap;
    unsigned char* sF = reinterpret_cast<unsigned char*>(sQa + nCap);
    // MULTI: shadow delta of the ring row above (side 0) / below (side 1) when that row belongs to a neighbouring rank
    T* sH = reinterpret_cast<T*>(smemRaw + (((size_t)(nRing * 7 + nCap * 5) * sizeof(T) + nCap + 15) / 16) * 16);   // [2 sides][3 planes][tw]
    const bool shadowUp = MULTI && M.toUp != nullptr && ty0 == D.y0, shadowDn = MULTI && M.toDn != nullptr && ty0 + thE == D.y1;
    if (MULTI) for (int c = tid; c < 6 * tw; c += NT) sH[c] = T(0);
    const Vec2<T>* pO = reinterpret_cast<const Vec2<T>*>(p);
    for (int c = tid; c < (thE + 2) * pw; c += NT) {              // ring-included planes; cells outside the image are zero
        const int lr = c / pw, lc = c - lr * pw, gx = tx0 - 1 + lc, gy = ty0 - 1 + lr;
        T px = T(0), py = T(0), pa = T(0); Vec4<T> g{ T(0), T(0), T(0), T(0) };
        if (lc <= twE + 1 && gx >= 0 && gx < W && gy >= 0 && gy < H) { const long j = (long)gy * W + gx; const Vec2<T> po = pO[j]; px = po.x; py = po.y; pa = p[D.offA + j]; g = D.geom[j]; }
        sPx[c] = px; sPy[c] = py; sPa[c] = pa; sGc[c] = g.x; sGs[c] = g.y; sGx[c] = g.z; sGy[c] = g.w;
    }
    T zx[KMAX], zy[KMAX], za[KMAX], ax[KMAX], ay[KMAX], aa[KMAX], dx[KMAX], dy[KMAX], da[KMAX];
    int ccOf[KMAX]; long giOf[KMAX];                              // ring-plane cell and global pixel index of the thread's pixels (-1: none)
    bool edge[KMAX];                                              // on the tile's perimeter: its z is read by a neighbouring tile
    bool pushUp[KMAX], pushDn[KMAX];                              // MULTI: first / last owned row with a neighbouring rank beyond it
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int c = tid + q * NT;
        zx[q] = zy[q] = za[q] = ax[q] = ay[q] = aa[q] = dx[q] = dy[q] = da[q] = T(0); edge[q] = false; ccOf[q] = -1; giOf[q] = -1; pushUp[q] = pushDn[q] = false;
        if (c < nCap) {
            const int ly = c / tw, lx = c - ly * tw;
            if (lx < twE && ly < thE) {
                const long i = (long)(ty0 + ly) * W + tx0 + lx;
                ccOf[q] = (ly + 1) * pw + lx + 1; giOf[q] = i;
                const Vec2<T> ro = reinterpret_cast<const Vec2<T>*>(r)[i]; sRx[c] = ro.x; sRy[c] = ro.y; sRa[c] = r[D.offA + i];
                sQo[c] = reinterpret_cast<const Vec2<T>*>(pre)[i].x; sQa[c] = pre[D.offA + i];
                sF[c] = D.flags[i];
                const Vec2<T> dO = reinterpret_cast<const Vec2<T>*>(delta)[i]; dx[q] = dO.x; dy[q] = dO.y; da[q] = delta[D.offA + i];
                edge[q] = (ly == 0 || ly == thE - 1 || lx == 0 || lx == twE - 1);
                if (MULTI) { pushUp[q] = M.toUp != nullptr && ty0 + ly == D.y0; pushDn[q] = M.toDn != nullptr && ty0 + ly == D.y1 - 1; }
            }
        }
    }
    __syncthreads();
    unsigned long long k = k0;
    double dRz, dRzPrev = 0.0, dPAp = 0.0;
    if (MULTI) {   // r.z of PCGInit1 was all-reduced through the mailboxes
        __shared__ double s_rz0;
        if (tid == 0) s_rz0 = loadScalarGlobal(ws, kRz0);
        __syncthreads();
        dRz = s_rz0;
    } else dRz = ws.scalars[kRz0];
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    const int nPerim = 2 * (twE + thE);
    for (int l = 0; l < L; ++l) {
        if (l > 0) {
            T beta = T(0);
            { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
#pragma unroll
            for (int q = 0; q < KMAX; ++q)
                if (ccOf[q] >= 0) { const int cc = ccOf[q]; sPx[cc] = zx[q] + beta * sPx[cc]; sPy[cc] = zy[q] + beta * sPy[cc]; sPa[cc] = za[q] + beta * sPa[cc]; }
            for (int c = tid; c < nPerim; c += NT) {              // ring cells: top row, bottom row, left column, right column
                int lr, lc;
                if (c < twE) { lr = 0; lc = c + 1; }
                else if (c < 2 * twE) { lr = thE + 1; lc = c - twE + 1; }
                else if (c < 2 * twE + thE) { lr = c - 2 * twE + 1; lc = 0; }
                else { lr = c - 2 * twE - thE + 1; lc = twE + 1; }
                const int gx = tx0 - 1 + lc, gy = ty0 - 1 + lr;
                if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                    const long j = (long)gy * W + gx;
                    Vec2<T> zo; T zaa;
                    const unsigned long long* ll = !MULTI ? nullptr : ((gy == D.y0 - 1 && M.fromUp) ? M.fromUp : ((gy == D.y1 && M.fromDn) ? M.fromDn : nullptr));
                    if (ll) {   // the row belongs to the neighbouring rank: its z arrives as LL words carrying this phase's sequence number
                        const unsigned long long* w = ll + (size_t)3 * gx * LLWords<T>::N;
                        llLoad(w, (unsigned)k, zo.x); llLoad(w + LLWords<T>::N, (unsigned)k, zo.y); llLoad(w + 2 * LLWords<T>::N, (unsigned)k, zaa);
                    } else { zo = ldcgV(reinterpret_cast<const Vec2<T>*>(z) + j); zaa = ldcgT(z + D.offA + j); }
                    const int cc = lr * pw + lc;
                    sPx[cc] = zo.x + beta * sPx[cc]; sPy[cc] = zo.y + beta * sPy[cc]; sPa[cc] = zaa + beta * sPa[cc];
                }
            }
            __syncthreads();
        }
        T acc = T(0);
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            T ox = T(0), oy = T(0), oa = T(0);
            if (ccOf[q] >= 0) {
                const int cc = ccOf[q];
                const unsigned fl = sF[tid + q * NT];
                if (fl & 32u) {
                    const T pix = sPx[cc], piy = sPy[cc], pai = sPa[cc];
                    const T gc = sGc[cc], gs = sGs[cc], gux = sGx[cc], guy = sGy[cc];
                    T sx = T(0), sy = T(0), sa = T(0);
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        if (fl & (1u << nb)) {
                            const int jj = cc + (nb == 0 ? 1 : (nb == 1 ? -1 : (nb == 2 ? pw : -pw)));
                            const T jc = sGc[jj], js = sGs[jj];
                            const T ddx = gux - sGx[jj], ddy = guy - sGy[jj];
                            const T eix = -gs * ddx - gc * ddy, eiy = gc * ddx - gs * ddy;
                            const T ejx = -js * ddx - jc * ddy, ejy = jc * ddx - js * ddy;
                            const T paj = sPa[jj];
                            const T px = pix - sPx[jj], py = piy - sPy[jj];
                            sx += T(2) * px - eix * pai - ejx * paj;
                            sy += T(2) * py - eiy * pai - ejy * paj;
                            sa += (eix * eix + eiy * eiy) * pai - (eix * px + eiy * py);
                        }
                    }
                    ox = w2 * sx; oy = w2 * sy; oa = w2 * sa;
                    if (fl & 16u) { ox += wf2 * pix; oy += wf2 * piy; }
                    acc += pix * ox + piy * oy + pai * oa;
                }
            }
            ax[q] = ox; ay[q] = oy; aa[q] = oa;
        }
        dPAp = MULTI ? gridAllReduceLLMulti<NT>((double)acc, llWords, arrivals, ++k, false, ws.comm, kLoopMailboxSlot)
                     : gridAllReduceLL<NT>((double)acc, llWords, arrivals, ++k, false);
        T alpha = T(0);
        { const T num = (T)dRz, den = (T)dPAp; if (den > T(0)) alpha = num / den; }
        if (MULTI) {   // shadow delta of the neighbour's row: same alpha, same p (the ring cells) => bitwise the neighbour's own delta
            if (shadowUp) for (int c = tid; c < twE; c += NT) { const int cc = c + 1; sH[c] += alpha * sPx[cc]; sH[tw + c] += alpha * sPy[cc]; sH[2 * tw + c] += alpha * sPa[cc]; }
            if (shadowDn) for (int c = tid; c < twE; c += NT) { const int cc = (thE + 1) * pw + c + 1; sH[3 * tw + c] += alpha * sPx[cc]; sH[4 * tw + c] += alpha * sPy[cc]; sH[5 * tw + c] += alpha * sPa[cc]; }
        }
        acc = T(0);
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            if (ccOf[q] >= 0) {
                const int cc = ccOf[q], c = tid + q * NT;
                dx[q] = dx[q] + alpha * sPx[cc]; dy[q] = dy[q] + alpha * sPy[cc]; da[q] = da[q] + alpha * sPa[cc];
                const T rx = sRx[c] - alpha * ax[q], ry = sRy[c] - alpha * ay[q], ra = sRa[c] - alpha * aa[q];
                sRx[c] = rx; sRy[c] = ry; sRa[c] = ra;
                const T qo = sQo[c];
                zx[q] = qo * rx; zy[q] = qo * ry; za[q] = sQa[c] * ra;
                acc += zx[q] * rx; acc += zy[q] * ry; acc += za[q] * ra;
                if (edge[q]) { stcgV(reinterpret_cast<Vec2<T>*>(z) + giOf[q], Vec2<T>{ zx[q], zy[q] }); stcgT(z + D.offA + giOf[q], za[q]); }
                if (MULTI && (pushUp[q] || pushDn[q])) {   // my edge row is the neighbouring rank's halo row: LL words over NVLink, no fence
                    const size_t x = (size_t)(giOf[q] % W);
                    const unsigned seq = (unsigned)(k + 1ull);   // = k after the reduction below: what the reader's phase A will expect
                    if (pushUp[q]) { unsigned long long* w = M.toUp + 3 * x * LLWords<T>::N; llStore(w, zx[q], seq); llStore(w + LLWords<T>::N, zy[q], seq); llStore(w + 2 * LLWords<T>::N, za[q], seq); }
                    if (pushDn[q]) { unsigned long long* w = M.toDn + 3 * x * LLWords<T>::N; llStore(w, zx[q], seq); llStore(w + LLWords<T>::N, zy[q], seq); llStore(w + 2 * LLWords<T>::N, za[q], seq); }
                }
            }
        }
        dRzPrev = dRz;
        dRz = MULTI ? gridAllReduceLLMulti<NT>((double)acc, llWords, arrivals, ++k, true, ws.comm, kLoopMailboxSlot)
                    : gridAllReduceLL<NT>((double)acc, llWords, arrivals, ++k, true);
    }
    T beta = T(0);
    if (L > 0) { const T rn = (T)dRz, ro = (T)dRzPrev; if (ro > T(0)) beta = rn / ro; }
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        if (ccOf[q] >= 0) {
            const int cc = ccOf[q], c = tid + q * NT; const long i = giOf[q];
            T px = sPx[cc], py = sPy[cc], pa = sPa[cc];
            if (L > 0) { px = zx[q] + beta * px; py = zy[q] + beta * py; pa = za[q] + beta * pa; }
            reinterpret_cast<Vec2<T>*>(p)[i] = Vec2<T>{ px, py }; p[D.offA + i] = pa;
            reinterpret_cast<Vec2<T>*>(delta)[i] = Vec2<T>{ dx[q], dy[q] }; delta[D.offA + i] = da[q];
            reinterpret_cast<Vec2<T>*>(r)[i] = Vec2<T>{ sRx[c], sRy[c] }; r[D.offA + i] = sRa[c];
            reinterpret_cast<Vec2<T>*>(z)[i] = Vec2<T>{ zx[q], zy[q] }; z[D.offA + i] = za[q];
            reinterpret_cast<Vec2<T>*>(Ap)[i] = Vec2<T>{ ax[q], ay[q] }; Ap[D.offA + i] = aa[q];
        }
    }
    if (MULTI) {   // delta of the neighbour's boundary row, so that the update of X on my halo rows matches the owner's
        if (shadowUp) for (int c = tid; c < twE; c += NT) { const long j = (long)(D.y0 - 1) * W + tx0 + c; reinterpret_cast<Vec2<T>*>(delta)[j] = Vec2<T>{ sH[c], sH[tw + c] }; delta[D.offA + j] = sH[2 * tw + c]; }
        if (shadowDn) for (int c = tid; c < twE; c += NT) { const long j = (long)D.y1 * W + tx0 + c; reinterpret_cast<Vec2<T>*>(delta)[j] = Vec2<T>{ sH[3 * tw + c], sH[4 * tw + c] }; delta[D.offA + j] = sH[5 * tw + c]; }
    }
    // single GPU: leave the scalars as the per-step loop does.  MULTI: the per-step loop keeps them in the mailboxes with their own
    // epochs; this kernel's sequence lives in a spare slot, so kRz0/kRz1/kPAp keep the values PCGInit1 published (nothing reads them
    // before the next PCGInit1 except the debugging read-back)
    if (!MULTI && L > 0 && blockIdx.x == 0 && tid == 0) { ws.scalars[kPAp] = dPAp; ws.scalars[L & 1] = dRz; ws.scalars[(L + 1) & 1] = dRzPrev; }
}

// ------------------------------------------------------------------------------------ whole PCG loop, state on chip, ONE reduction per step
// The two dependent all-reduces of a PCG step (p.Ap before alpha, r.z before beta) are what the collective loop kernel above spends its
// time in: 19.6 us per step on 8 GPUs, ~4 us of it arithmetic.  This kernel runs the SAME Jacobi-PCG in the Chronopoulos-Gear
// arrangement (the "single allreduce of the two PCG scalars" of BASELINE.json's north star): the matvec is applied to z instead of p,
//     w = A z,  gamma = r.z,  delta = w.z            <- ONE all-reduce of two scalars per step
//     beta = gamma / gamma_old,   alpha = gamma / (delta - beta gamma / alpha_old)      (= r.z / p.Ap, since p.Ap = delta - beta gamma / alpha_old)
//     p = z + beta p,  s = w + beta s (= A p),  delta_x += alpha p,  r -= alpha s,  z = pre r
// In exact arithmetic these are the reference's iterates (solverGPUGaussNewton.t:421-550, same guards: a non-positive denominator gives
// alpha = 0, a non-positive gamma_old gives beta = 0); in floating point A p is carried by a recurrence instead of being applied, so the
// iterates differ in the last bits (tests: state after k steps vs the oracle to the usual 3e-4; a 400-step fp32 prototype against fp64
// loses nothing over the two-reduction form).  Tile state: z (with ring), geometry (with ring), r, pre, flags in shared memory;
// p, s, delta_x, w of a thread's pixels in registers (512 threads x up to 8 pixels).  Tile edges of z go to the neighbouring tiles as LL
// data words (edgeLL: one row of words per tile side; llStore / llLoad), and to a neighbouring RANK through the peer scratch rows as in
// k_iwPcgLoopTile -- pure dataflow: the kernel contains no fence and no barrier besides the all-reduce itself.
template <class T, int KMAX, bool MULTI>
__global__ void __launch_bounds__(512, 1) k_iwCgLoopTile(IWData<T> D, ReduceWorkspace ws, unsigned long long* llWords, unsigned long long* arrivals, unsigned long long* edgeLL,
                                                         unsigned long long k0, int L, int tw, int th, int tilesAcross, int tilesDown, IWTileMulti<T> M,
                                                         T* delta, T* r, T* z, T* p, T* Ap, const T* __restrict__ pre) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    constexpr int NT = 512, WN = LLWords<T>::N;
    const int W = D.W, H = D.H, tid = threadIdx.x;
    const int tcol = blockIdx.x % tilesAcross, trow = blockIdx.x / tilesAcross;
    const int tx0 = tcol * tw, ty0 = D.y0 + trow * th;
    const int twE = min(tw, W - tx0), thE = min(th, D.y1 - ty0);
    const int pw = tw + 2;
    const int nRing = (th + 2) * pw, nCap = tw * th;
    T* sZx = reinterpret_cast<T*>(smemRaw); T* sZy = sZx + nRing; T* sZa = sZy + nRing;
    T* sGc = sZa + nRing; T* sGs = sGc + nRing; T* sGx = sGs + nRing; T* sGy = sGx + nRing;
    T* sRx = sGy + nRing; T* sRy = sRx + nCap; T* sRa = sRy + nCap; T* sQo = sRa + nCap; T* sQa = sQo + nCap;
    unsigned char* sF = reinterpret_cast<unsigned char*>(sQa + nCap);
    // LL edge rows of a tile: [top tw][bottom tw][left th][right th], 3 scalars per cell
    const size_t edgeStride = (size_t)2 * (tw + th) * 3 * WN;
    auto edgeOf = [&](int tile, int side) -> unsigned long long* {
        return edgeLL + (size_t)tile * edgeStride + (size_t)(side == 0 ? 0 : (side == 1 ? tw : (side == 2 ? 2 * tw : 2 * tw + th))) * 3 * WN;
    };
    // prologue: z0 = the initial direction p = pre r (PCGInit1), ring cells from the p vector (its halo rows were exchanged by the caller)
    const Vec2<T>* pO = reinterpret_cast<const Vec2<T>*>(p);
    for (int c = tid; c < (th + 2) * pw; c += NT) {
        const int lr = c / pw, lc = c - lr * pw, gx = tx0 - 1 + lc, gy = ty0 - 1 + lr;
        T zx = T(0), zy = T(0), za = T(0); Vec4<T> g{ T(0), T(0), T(0), T(0) };
        if (lr <= thE + 1 && lc <= twE + 1 && gx >= 0 && gx < W && gy >= 0 && gy < H) { const long j = (long)gy * W + gx; const Vec2<T> po = pO[j]; zx = po.x; zy = po.y; za = p[D.offA + j]; g = D.geom[j]; }
        sZx[c] = zx; sZy[c] = zy; sZa[c] = za; sGc[c] = g.x; sGs[c] = g.y; sGx[c] = g.z; sGy[c] = g.w;
    }
    T px[KMAX], py[KMAX], pa[KMAX], sx[KMAX], sy[KMAX], sa[KMAX], dx[KMAX], dy[KMAX], da[KMAX], wx[KMAX], wy[KMAX], wa[KMAX];
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        px[q] = py[q] = pa[q] = sx[q] = sy[q] = sa[q] = dx[q] = dy[q] = da[q] = wx[q] = wy[q] = wa[q] = T(0);
        const int c = tid + q * NT;
        if (c < nCap) {
            const int ly = c / tw, lx = c - ly * tw;
            sF[c] = 0; sRx[c] = sRy[c] = sRa[c] = sQo[c] = sQa[c] = T(0);
            if (lx < twE && ly < thE) {
                const long i = (long)(ty0 + ly) * W + tx0 + lx;
                const Vec2<T> ro = reinterpret_cast<const Vec2<T>*>(r)[i]; sRx[c] = ro.x; sRy[c] = ro.y; sRa[c] = r[D.offA + i];
                sQo[c] = reinterpret_cast<const Vec2<T>*>(pre)[i].x; sQa[c] = pre[D.offA + i];
                sF[c] = D.flags[i];
                const Vec2<T> dO = reinterpret_cast<const Vec2<T>*>(delta)[i]; dx[q] = dO.x; dy[q] = dO.y; da[q] = delta[D.offA + i];
            }
        }
    }
    __syncthreads();
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    // w = A z for the thread's pixels, and the two partial dot products
    auto matvec = [&](T& accG, T& accD) {
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            const int c = tid + q * NT;
            T ox = T(0), oy = T(0), oa = T(0);
            if (c < nCap) {
                const int ly = c / tw, lx = c - ly * tw, cc = (ly + 1) * pw + lx + 1;
                const unsigned fl = sF[c];
                if (fl & 32u) {
                    const T zix = sZx[cc], ziy = sZy[cc], zai = sZa[cc];
                    const T gc = sGc[cc], gs = sGs[cc], gux = sGx[cc], guy = sGy[cc];
                    T ax = T(0), ay = T(0), aa = T(0);
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) {
                        if (fl & (1u << nb)) {
                            const int jj = cc + (nb == 0 ? 1 : (nb == 1 ? -1 : (nb == 2 ? pw : -pw)));
                            const T jc = sGc[jj], js = sGs[jj];
                            const T ddx = gux - sGx[jj], ddy = guy - sGy[jj];
                            const T eix = -gs * ddx - gc * ddy, eiy = gc * ddx - gs * ddy;
                            const T ejx = -js * ddx - jc * ddy, ejy = jc * ddx - js * ddy;
                            const T zaj = sZa[jj];
                            const T qx = zix - sZx[jj], qy = ziy - sZy[jj];
                            ax += T(2) * qx - eix * zai - ejx * zaj;
                            ay += T(2) * qy - eiy * zai - ejy * zaj;
                            aa += (eix * eix + eiy * eiy) * zai - (eix * qx + eiy * qy);
                        }
                    }
                    ox = w2 * ax; oy = w2 * ay; oa = w2 * aa;
                    if (fl & 16u) { ox += wf2 * zix; oy += wf2 * ziy; }
                    accD += zix * ox + ziy * oy + zai * oa;
                    accG += sRx[c] * zix; accG += sRy[c] * ziy; accG += sRa[c] * zai;
                }
            }
            wx[q] = ox; wy[q] = oy; wa[q] = oa;
        }
    };
    unsigned long long k = k0;
    T accG = T(0), accD = T(0);
    matvec(accG, accD);
    double dG, dD;
    gridAllReduce2LL<NT, MULTI>((double)accG, (double)accD, llWords, arrivals, ++k, ws.comm, kLoopMailboxSlot, dG, dD);
    T gam = (T)dG, alpha = T(0), beta = T(0);
    { const T den = (T)dD; if (den > T(0)) alpha = gam / den; }
    for (int it = 0; it < L; ++it) {
        const bool more = it + 1 < L;
        const unsigned seq = (unsigned)(k + 1ull);     // = the k of the all-reduce that follows: tags z of iteration it + 1
        // ---------------- updates (PCGStep3 of the previous step, PCGStep2 of this one), then z of the next step goes to shared memory and to the neighbours
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
            const int c = tid + q * NT;
            if (c < nCap) {
                const int ly = c / tw, lx = c - ly * tw;
                if (lx < twE && ly < thE) {
                    const int cc = (ly + 1) * pw + lx + 1;
                    px[q] = sZx[cc] + beta * px[q]; py[q] = sZy[cc] + beta * py[q]; pa[q] = sZa[cc] + beta * pa[q];
                    sx[q] = wx[q] + beta * sx[q]; sy[q] = wy[q] + beta * sy[q]; sa[q] = wa[q] + beta * sa[q];
                    dx[q] = dx[q] + alpha * px[q]; dy[q] = dy[q] + alpha * py[q]; da[q] = da[q] + alpha * pa[q];
                    const T rx = sRx[c] - alpha * sx[q], ry = sRy[c] - alpha * sy[q], ra = sRa[c] - alpha * sa[q];
                    sRx[c] = rx; sRy[c] = ry; sRa[c] = ra;
                    const T qo = sQo[c];
                    const T zx = qo * rx, zy = qo * ry, za = sQa[c] * ra;
                    sZx[cc] = zx; sZy[cc] = zy; sZa[cc] = za;
                    if (more) {
                        if (ly == 0) {
                            if (trow > 0) { unsigned long long* e = edgeOf(blockIdx.x, 0) + (size_t)lx * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                            else if (MULTI && M.toUp) { unsigned long long* e = M.toUp + (size_t)(tx0 + lx) * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                        }
                        if (ly == thE - 1) {
                            if (trow + 1 < tilesDown) { unsigned long long* e = edgeOf(blockIdx.x, 1) + (size_t)lx * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                            else if (MULTI && M.toDn) { unsigned long long* e = M.toDn + (size_t)(tx0 + lx) * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                        }
                        if (lx == 0 && tcol > 0) { unsigned long long* e = edgeOf(blockIdx.x, 2) + (size_t)ly * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                        if (lx == twE - 1 && tcol + 1 < tilesAcross) { unsigned long long* e = edgeOf(blockIdx.x, 3) + (size_t)ly * 3 * WN; llStore(e, zx, seq); llStore(e + WN, zy, seq); llStore(e + 2 * WN, za, seq); }
                    }
                }
            }
        }
        if (!more) break;
        // ---------------- ring cells: the neighbours' z of the same step (LL words: the poll IS the synchronisation)
        for (int c = tid; c < 2 * (twE + thE); c += NT) {
            int lr, lc; const unsigned long long* e = nullptr;
            if (c < twE) {                     // row above
                lr = 0; lc = c + 1;
                if (trow > 0) e = edgeOf(blockIdx.x - tilesAcross, 1) + (size_t)c * 3 * WN;
                else if (MULTI && M.fromUp) e = M.fromUp + (size_t)(tx0 + c) * 3 * WN;
            } else if (c < 2 * twE) {          // row below
                const int i = c - twE; lr = thE + 1; lc = i + 1;
                if (trow + 1 < tilesDown) e = edgeOf(blockIdx.x + tilesAcross, 0) + (size_t)i * 3 * WN;
                else if (MULTI && M.fromDn) e = M.fromDn + (size_t)(tx0 + i) * 3 * WN;
            } else if (c < 2 * twE + thE) {    // column to the left
                const int i = c - 2 * twE; lr = i + 1; lc = 0;
                if (tcol > 0) e = edgeOf(blockIdx.x - 1, 3) + (size_t)i * 3 * WN;
            } else {                           // column to the right
                const int i = c - 2 * twE - thE; lr = i + 1; lc = twE + 1;
                if (tcol + 1 < tilesAcross) e = edgeOf(blockIdx.x + 1, 2) + (size_t)i * 3 * WN;
            }
            if (e) { const int cc = lr * pw + lc; T a, b, cval; llLoad(e, seq, a); llLoad(e + WN, seq, b); llLoad(e + 2 * WN, seq, cval); sZx[cc] = a; sZy[cc] = b; sZa[cc] = cval; }
        }
        __syncthreads();
        // ---------------- w = A z, gamma, delta: the step's only all-reduce
        accG = T(0); accD = T(0);
        matvec(accG, accD);
        gridAllReduce2LL<NT, MULTI>((double)accG, (double)accD, llWords, arrivals, ++k, ws.comm, kLoopMailboxSlot, dG, dD);
        const T gNew = (T)dG, dNew = (T)dD;
        beta = T(0); if (gam > T(0)) beta = gNew / gam;                                  // :541-546
        T den = dNew; if (alpha != T(0)) den = dNew - beta * gNew / alpha;               // = p.Ap of the coming step
        alpha = T(0); if (den > T(0)) alpha = gNew / den;                                // :456-461
        gam = gNew;
    }
    // the state goes back to the plan's vectors: delta, r, z; p and Ap (= s) as they were used by the last update
#pragma unroll
    for (int q = 0; q < KMAX; ++q) {
        const int c = tid + q * NT;
        if (c < nCap) {
            const int ly = c / tw, lx = c - ly * tw;
            if (lx < twE && ly < thE) {
                const int cc = (ly + 1) * pw + lx + 1; const long i = (long)(ty0 + ly) * W + tx0 + lx;
                reinterpret_cast<Vec2<T>*>(p)[i] = Vec2<T>{ px[q], py[q] }; p[D.offA + i] = pa[q];
                reinterpret_cast<Vec2<T>*>(delta)[i] = Vec2<T>{ dx[q], dy[q] }; delta[D.offA + i] = da[q];
                reinterpret_cast<Vec2<T>*>(r)[i] = Vec2<T>{ sRx[c], sRy[c] }; r[D.offA + i] = sRa[c];
                reinterpret_cast<Vec2<T>*>(z)[i] = Vec2<T>{ sZx[cc], sZy[cc] }; z[D.offA + i] = sZa[cc];
                reinterpret_cast<Vec2<T>*>(Ap)[i] = Vec2<T>{ sx[q], sy[q] }; Ap[D.offA + i] = sa[q];
            }
        }
    }
}

// ------------------------------------------------------------------------------------ computeCost
// solverGPUGaussNewton.t:580-592 + o.t:2375-2385: 1/2 sum r^2 over residuals centred at active pixels.
// MODEL: computeModelCost (:666-678, o.t:2174-2199): 1/2 sum (r + J delta)^2.
template <class T, bool MODEL>
__global__ void __launch_bounds__(kBlock) k_iwCost(IWData<T> D, ReduceWorkspace ws, int slot, const T* __restrict__ delta) {
    T acc = T(0);
    const Vec2<T>* dO = reinterpret_cast<const Vec2<T>*>(delta);
    const T* dA = MODEL ? delta + D.offA : nullptr;
    IW_TILE_LOOP(D) {
        const long i = (long)y * D.W + x;
        if (y >= D.y0 && y < D.y1 && D.Mask[i] == T(0)) {
            const Vec2<T> Oi = D.Offset[i], Ui = D.UrShape[i];
            T s, c; sincosT(D.Angle[i], &s, &c);
            Vec2<T> di{ T(0), T(0) }; T dai = T(0);
            if (MODEL) { di = dO[i]; dai = dA[i]; }
            T e = T(0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                long j;
                if (neighbour(k, x, y, D.W, D.H, &j) && D.Mask[j] == T(0)) {
                    const Vec2<T> Oj = D.Offset[j], Uj = D.UrShape[j];
                    const T dx = Ui.x - Uj.x, dy = Ui.y - Uj.y;
                    T rx = D.w * ((Oi.x - Oj.x) - (c * dx - s * dy)), ry = D.w * ((Oi.y - Oj.y) - (s * dx + c * dy));
                    if (MODEL) {
                        const Vec2<T> dj = dO[j];
                        const T ex = -s * dx - c * dy, ey = c * dx - s * dy;
                        rx += D.w * ((di.x - dj.x) - ex * dai); ry += D.w * ((di.y - dj.y) - ey * dai);
                    }
                    e += rx * rx + ry * ry;
                }
            }
            const Vec2<T> C = D.Constraints[i];
            if (C.x >= T(0) && C.y >= T(0)) {
                T fx = D.wf * (Oi.x - C.x), fy = D.wf * (Oi.y - C.y);
                if (MODEL) { fx += D.wf * di.x; fy += D.wf * di.y; }
                e += fx * fx + fy * fy;
            }
            acc += T(0.5) * e;
        }
    }
    gridReduceTo<kBlock>((double)acc, ws, slot);
}

// ------------------------------------------------------------------------------------ PCGComputeCtC
// :616-622 + o.t:2255-2287: CtC = diag(J^T J) / radius, from the packed linearisation.
template <class T>
__global__ void __launch_bounds__(kBlock) k_iwCtC(IWData<T> D, T* __restrict__ CtC, T invRadius) {
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    IW_TILE_LOOP(D) {
        const long i = (long)y * D.W + x;
        const unsigned fl = D.flags[i];
        T dO = T(0), dA = T(0);
        if (fl & 32u) {
            const Vec4<T> gi = D.geom[i];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (fl & (1u << k)) {
                    long j; neighbour(k, x, y, D.W, D.H, &j);
                    const Vec4<T> gj = D.geom[j];
                    const T dx = gi.z - gj.z, dy = gi.w - gj.w;
                    const T ex = -gi.y * dx - gi.x * dy, ey = gi.x * dx - gi.y * dy;
                    dO += T(2) * w2 * invRadius;
                    dA += w2 * (ex * ex + ey * ey) * invRadius;
                }
            }
            if (fl & 16u) dO += wf2 * invRadius;
        }
        reinterpret_cast<Vec2<T>*>(CtC)[i] = Vec2<T>{ dO, dO }; CtC[D.offA + i] = dA;
    }
}

// ====================================================================================== host side
template <class T>
struct ImageWarpingOps : EnergyOps<T> {
    IWData<T> D{};
    float wFit = 0, wReg = 0;
    int gridInit = 0, gridJTJ = 0, gridJTJF = 0, gridCost = 0;
    // TMA path (fp32, W % 4 == 0): tensor maps are cached per (p vector, z vector) pair
    bool tmaOk = false; int tmaTilesX = 0, tmaTiles = 0, gridTma[2] = { 0, 0 }, gridTmaF[2] = { 0, 0 };   // [0]: multi-rank shape (2 stages), [1]: single GPU (IW_STAGES)
    std::map<std::pair<const void*, const void*>, IWMaps> mapCache;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    EncodeFn encode = nullptr;
    CUtensorMap map2D(const float* base, long innerFloats, long rows, long pitchFloats, int boxInner, int boxRows) {
        CUtensorMap m; cuuint64_t dims[2] = { (cuuint64_t)innerFloats, (cuuint64_t)rows }; cuuint64_t strides[1] = { (cuuint64_t)pitchFloats * 4 };
        cuuint32_t box[2] = { (cuuint32_t)boxInner, (cuuint32_t)boxRows }, es[2] = { 1, 1 };
        CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { fprintf(stderr, "opt_b200: cuTensorMapEncodeTiled(2D) failed: %d\n", (int)rc); exit(3); }
        return m;
    }
    CUtensorMap mapGeom() {
        CUtensorMap m; cuuint64_t dims[3] = { 4, (cuuint64_t)D.W, (cuuint64_t)D.H }; cuuint64_t strides[2] = { 16, (cuuint64_t)D.W * 16 };
        cuuint32_t box[3] = { 4, (cuuint32_t)(kTW + 2), (cuuint32_t)kRows }, es[3] = { 1, 1, 1 };
        CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)D.geom, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { fprintf(stderr, "opt_b200: cuTensorMapEncodeTiled(3D) failed: %d\n", (int)rc); exit(3); }
        return m;
    }
    CUtensorMap mapFlags() {
        CUtensorMap m; cuuint64_t dims[2] = { (cuuint64_t)D.W, (cuuint64_t)D.H }; cuuint64_t strides[1] = { (cuuint64_t)D.W };
        cuuint32_t box[2] = { (cuuint32_t)kTW, (cuuint32_t)kTH }, es[2] = { 1, 1 };
        CUresult rc = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)D.flags, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { fprintf(stderr, "opt_b200: cuTensorMapEncodeTiled(flags) failed: %d\n", (int)rc); exit(3); }
        return m;
    }
    const IWMaps& mapsFor(const float* pIn, const float* z) {
        auto key = std::make_pair((const void*)pIn, (const void*)z);
        auto it = mapCache.find(key);
        if (it != mapCache.end()) return it->second;
        IWMaps m;
        m.pO = map2D(pIn, 2L * D.W, D.H, 2L * D.W, kRowO, kRows);
        m.pA = map2D(pIn + D.offA, D.W, D.H, D.W, kRowA, kRows);
        const float* zz = z ? z : pIn;
        m.zO = map2D(zz, 2L * D.W, D.H, 2L * D.W, kRowO, kRows);
        m.zA = map2D(zz + D.offA, D.W, D.H, D.W, kRowA, kRows);
        m.geom = mapGeom();
        m.flags = mapFlags();
        return mapCache.emplace(key, m).first->second;
    }
    void setupTma(int sms) {
        if (!std::is_same<T, float>::value || D.W % 16 != 0 || getenv("OPTB_NO_TMA")) return;   // 16-byte row pitch rule of the flag-byte map
        void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) return;
        encode = (EncodeFn)fn;
        tmaTilesX = (D.W + kTW - 1) / kTW; tmaTiles = tmaTilesX * ((D.H + kTH - 1) / kTH);
        auto prep = [&](const void* f, int bytes) {
            OPT_CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, f, kBlock, bytes); if (occ < 1) occ = 1;
            long g = (long)sms * occ; if (g > tmaTiles) g = tmaTiles; if (g > kMaxPartials) g = kMaxPartials; return (int)(g < 1 ? 1 : g);
        };
        gridTma[0] = prep((const void*)k_iwApplyJTJTma<false, false, kStagesMulti>, kStagesMulti * kStageBytesPlain);
        prep((const void*)k_iwApplyJTJTma<false, true, kStagesMulti>, kStagesMulti * kStageBytesPlain);
        gridTmaF[0] = prep((const void*)k_iwApplyJTJTma<true, false, kStagesMulti>, kStagesMulti * kStageBytesFused);
        prep((const void*)k_iwApplyJTJTma<true, true, kStagesMulti>, kStagesMulti * kStageBytesFused);
        gridTma[1] = prep((const void*)k_iwApplyJTJTma<false, false, kStagesSingle>, kStagesSingle * kStageBytesPlain);
        prep((const void*)k_iwApplyJTJTma<false, true, kStagesSingle>, kStagesSingle * kStageBytesPlain);
        gridTmaF[1] = prep((const void*)k_iwApplyJTJTma<true, false, kStagesSingle>, kStagesSingle * kStageBytesFused);
        prep((const void*)k_iwApplyJTJTma<true, true, kStagesSingle>, kStagesSingle * kStageBytesFused);
        tmaOk = true;
    }
    ImageWarpingOps(const unsigned* dims) {
        D.W = (int)dims[0]; D.H = (int)dims[1]; D.y0 = 0; D.y1 = D.H;
        D.tilesX = (D.W + kTileW - 1) / kTileW; D.numTiles = D.tilesX * ((D.H + kTileH - 1) / kTileH);
        const long N = (long)D.W * D.H;
        this->unknowns.resize(2);
        this->unknowns[0].elements = N; this->unknowns[0].channels = 2;   // Offset
        this->unknowns[1].elements = N; this->unknowns[1].channels = 1;   // Angle
        this->layoutUnknowns();
        D.offA = this->unknowns[1].offset;
        this->usePreconditioner = true;
        OPT_CUDA_CHECK(cudaMalloc((void**)&D.geom, N * sizeof(Vec4<T>)));
        OPT_CUDA_CHECK(cudaMalloc((void**)&D.flags, N));
        OPT_CUDA_CHECK(cudaMemset(D.flags, 0, N));
        int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        auto gridOf = [&](const void* fn) {
            int occ = 1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBlock, 0);
            if (occ < 1) occ = 1;
            long g = (long)sms * occ; if (g > D.numTiles) g = D.numTiles; if (g > kMaxPartials) g = kMaxPartials; if (g < 1) g = 1;
            return (int)g;
        };
        gridInit = gridOf((const void*)k_iwInit<T>);
        gridJTJ = gridOf((const void*)k_iwApplyJTJ<T, false, false>);
        gridJTJF = gridOf((const void*)k_iwApplyJTJ<T, true, false>);
        gridCost = gridOf((const void*)k_iwCost<T, false>);
        setupTma(sms);
    }
    ~ImageWarpingOps() override { cudaFree(D.geom); cudaFree(D.flags); if (loopSync) cudaFree(loopSync); if (llWords) cudaFree(llWords); if (tileWords) cudaFree(tileWords); if (cgWords) cudaFree(cgWords); if (edgeLL) cudaFree(edgeLL); }
    const char* name() const override { return "image_warping"; }
    bool capturable() const override { return true; }
    unsigned long long captureKey() const override { return EnergyOps<T>::hashValue(D); }
    int upperRows = 0, lowerRows = 0;
    bool setSlab(int ownedBegin, int ownedEnd, int up, int dn) override {
        if (ownedBegin < 0 || ownedEnd > D.H || ownedBegin >= ownedEnd || D.W % 4 != 0) return false;
        D.y0 = ownedBegin; D.y1 = ownedEnd; upperRows = up; lowerRows = dn;
        return true;
    }
    static long alignUp64(long v) { return (v + 63) / 64 * 64; }
    std::vector<typename EnergyOps<T>::HaloSegment> haloSegments() const override {
        std::vector<typename EnergyOps<T>::HaloSegment> s;
        const long W = D.W;
        if (upperRows > 0) {   // my first owned row -> last local row of the upper neighbour
            s.push_back({ 2 * W * D.y0, 2 * W * (upperRows - 1), 2 * W, 0 });
            s.push_back({ D.offA + W * D.y0, alignUp64(2 * W * upperRows) + W * (upperRows - 1), W, 0 });
        }
        if (lowerRows > 0) {   // my last owned row -> local row 0 of the lower neighbour
            s.push_back({ 2 * W * (D.y1 - 1), 0, 2 * W, 1 });
            s.push_back({ D.offA + W * (D.y1 - 1), alignUp64(2 * W * lowerRows), W, 1 });
        }
        return s;
    }
    std::vector<std::pair<long, long>> ownedScalarRanges() const override {
        const long W = D.W;
        return { { 2 * W * D.y0, 2 * W * D.y1 }, { D.offA + W * D.y0, D.offA + W * D.y1 } };
    }
    void bind(void** q) override {   // positions: image_warping.t:2-8
        this->unknowns[0].user = q[0]; this->unknowns[1].user = q[1];
        D.Offset = (const Vec2<T>*)q[0]; D.Angle = (const T*)q[1]; D.UrShape = (const Vec2<T>*)q[2];
        D.Constraints = (const Vec2<T>*)q[3]; D.Mask = (const T*)q[4];
        wFit = *(const float*)q[5]; wReg = *(const float*)q[6];
        D.wf = (T)wFit; D.w = (T)wReg;
    }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        k_iwInit<T><<<gridInit, kBlock, 0, c.stream>>>(D, c.ws, rzSlot, delta, r, pre, p);
    }
    bool inKernelHalo() const override { return tmaOk; }
    // state-on-chip variant: one slab of R full rows per SM
    int resR = 0, resGrid = 0, resK = 0; size_t resSmem = 0; int resChecked = 0; unsigned long long* llWords = nullptr; unsigned long long resSeq = 0;
    template <int KMAX>
    bool launchResidentK(const LaunchCtx& c, int L, T* delta, T* r, T* z, T* p, T* Ap, const T* pre) {
        const void* fn = (const void*)k_iwPcgLoopResident<T, KMAX>;
        if (resChecked != KMAX) {
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)resSmem) != cudaSuccess) { cudaGetLastError(); return false; }
            int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 1024, resSmem);
            int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if ((long)occ * sms < resGrid) { resR = -1; return false; }   // does not fit this device: remember, do not ask again every step
            resChecked = KMAX;
        }
        ReduceWorkspace ws = c.ws;
        constexpr size_t llCount = 4 * (kMaxLLCtas + 1) + 1;   // two parities of (slots + result) words, then the arrival counter
        if (!llWords) { OPT_CUDA_CHECK(cudaMalloc((void**)&llWords, sizeof(unsigned long long) * llCount)); OPT_CUDA_CHECK(cudaMemset(llWords, 0, sizeof(unsigned long long) * llCount)); resSeq = 0; }
        unsigned long long* arrivals = llWords + 4 * (kMaxLLCtas + 1);
        unsigned long long k0 = resSeq; resSeq += 2ull * (unsigned long long)L;   // own sequence: the arrival counter counts gridDim per reduction
        void* args[] = { (void*)&D, (void*)&ws, (void*)&llWords, (void*)&arrivals, (void*)&k0, (void*)&L, (void*)&resR, (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&Ap, (void*)&pre };
        const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3((unsigned)resGrid), dim3(1024), args, resSmem, c.stream);
        if (le != cudaSuccess) {   // e.g. fewer SMs available than the device reports (MPS limits): nothing ran, use the per-step loop from now on
            cudaGetLastError(); resSeq = k0; resR = -1;
            return false;
        }
        return true;
    }
    // 2-D tile variant (opt-in, OPTB_RESIDENT_TILE=WxH; chosen automatically for the multi-GPU form): own sequence/arrival
    // state because its grid size differs
    int tileW = 0, tileH = 0, tileGrid = 0, tileK = 0, tileAcross = 0; size_t tileSmem = 0; unsigned long long* tileWords = nullptr; unsigned long long tileSeq = 0;
    bool tileMulti = false; const void* tileChecked = nullptr;
    template <int KMAX, bool MULTI>
    bool launchTileK(const LaunchCtx& c, int L, const IWTileMulti<T>& M, T* delta, T* r, T* z, T* p, T* Ap, const T* pre) {
        const void* fn = (const void*)k_iwPcgLoopTile<T, KMAX, MULTI>;
        if (tileChecked != fn) {   // once per kernel variant, not per launch
            if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tileSmem) != cudaSuccess) { cudaGetLastError(); return false; }
            int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 1024, tileSmem);
            int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if ((long)occ * sms < tileGrid) return false;
            tileChecked = fn;
        }
        constexpr size_t llCount = 4 * (kMaxLLCtas + 1) + 1;
        if (!tileWords) { OPT_CUDA_CHECK(cudaMalloc((void**)&tileWords, sizeof(unsigned long long) * llCount)); OPT_CUDA_CHECK(cudaMemset(tileWords, 0, sizeof(unsigned long long) * llCount)); tileSeq = 0; }
        unsigned long long* arrivals = tileWords + 4 * (kMaxLLCtas + 1);
        unsigned long long k0 = tileSeq; tileSeq += 2ull * (unsigned long long)L;
        ReduceWorkspace ws = c.ws;
        IWTileMulti<T> Mv = M;
        void* args[] = { (void*)&D, (void*)&ws, (void*)&tileWords, (void*)&arrivals, (void*)&k0, (void*)&L, (void*)&tileW, (void*)&tileH, (void*)&tileAcross, (void*)&Mv,
                         (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&Ap, (void*)&pre };
        const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3((unsigned)tileGrid), dim3(1024), args, tileSmem, c.stream);
        if (le != cudaSuccess) {
            if (MULTI) OPT_CUDA_CHECK(le);   // the other ranks are already inside the collective kernel: failing loudly is the only safe answer
            cudaGetLastError(); tileSeq = k0; tileW = -1;
            return false;
        }
        return true;
    }
    // single-reduction (Chronopoulos-Gear) form of the tile loop: OPT-IN (OPTB_CG_LOOP=1), single GPU and multi-rank.  The default loops
    // keep the reference's recurrences exactly (the collective two-reduction kernel is bitwise equal to the single-GPU on-chip loop); this
    // one is faster where the all-reduces dominate but carries A p by a recurrence: in fp32 its iterates drift from the reference's by
    // more than a summation-order change does (measured: cost after 40 un-converged PCG steps off by 1.3e-3 on a 128x96 case where the
    // two-reduction kernel is exact; 1e-6 on others) -- parity first, so it is not a default.
    unsigned long long* cgWords = nullptr; unsigned long long* edgeLL = nullptr; unsigned long long cgSeq = 0; size_t edgeWords = 0; const void* cgChecked = nullptr;
    bool useCg(bool) const { const char* e = getenv("OPTB_CG_LOOP"); return e && atoi(e) != 0; }
    int cgKmax() const { const long need = ((long)tileW * tileH + 511) / 512; return need <= 4 ? 4 : (need <= 7 ? 7 : 8); }
    template <int KMAX, bool MULTI> const void* cgFn() const { return (const void*)k_iwCgLoopTile<T, KMAX, MULTI>; }
    const void* cgFnFor(bool multi) const {
        const int km = cgKmax();
        if (multi) return km == 4 ? cgFn<4, true>() : (km == 7 ? cgFn<7, true>() : cgFn<8, true>());
        return km == 4 ? cgFn<4, false>() : (km == 7 ? cgFn<7, false>() : cgFn<8, false>());
    }
    bool cgPrepare(const void* fn) {   // attributes + co-residency, once per kernel variant
        if (cgChecked == fn) return true;
        const size_t smem = tileSmemFor(tileW, tileH, false);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
        int occ = 0, dev = 0, sms = 148; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 512, smem);
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if ((long)occ * sms < tileGrid) return false;
        cgChecked = fn;
        return true;
    }
    bool launchCg(const LaunchCtx& c, int L, bool multi, const IWTileMulti<T>& M, T* delta, T* r, T* z, T* p, T* Ap, const T* pre) {
        const void* fn = cgFnFor(multi);
        if (!cgPrepare(fn)) return false;
        constexpr size_t llCount = 8 * (kMaxLLCtas + 1) + 1;   // two parities of (slots + result) x 4 words, then the arrival counter
        const size_t needEdge = (size_t)tileGrid * 2 * (tileW + tileH) * 3 * LLWords<T>::N;
        if (!cgWords || needEdge != edgeWords) {   // (re)allocated zeroed: sequence numbers restart with the buffers
            if (cgWords) cudaFree(cgWords); if (edgeLL) cudaFree(edgeLL);
            OPT_CUDA_CHECK(cudaMalloc((void**)&cgWords, sizeof(unsigned long long) * llCount)); OPT_CUDA_CHECK(cudaMemset(cgWords, 0, sizeof(unsigned long long) * llCount));
            OPT_CUDA_CHECK(cudaMalloc((void**)&edgeLL, sizeof(unsigned long long) * std::max<size_t>(needEdge, 1))); OPT_CUDA_CHECK(cudaMemset(edgeLL, 0, sizeof(unsigned long long) * std::max<size_t>(needEdge, 1)));
            edgeWords = needEdge; cgSeq = 0;
        }
        unsigned long long* arrivals = cgWords + 8 * (kMaxLLCtas + 1);
        unsigned long long k0 = cgSeq; cgSeq += (unsigned long long)L;   // exactly L all-reduces per launch (the arrival counter counts gridDim per reduction)
        ReduceWorkspace ws = c.ws;
        IWTileMulti<T> Mv = M;
        const int tilesDown = tileGrid / tileAcross;
        const size_t smem = tileSmemFor(tileW, tileH, false);
        void* args[] = { (void*)&D, (void*)&ws, (void*)&cgWords, (void*)&arrivals, (void*)&edgeLL, (void*)&k0, (void*)&L, (void*)&tileW, (void*)&tileH, (void*)&tileAcross, (void*)&tilesDown,
                         (void*)&Mv, (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&Ap, (void*)&pre };
        const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3((unsigned)tileGrid), dim3(512), args, smem, c.stream);
        if (le != cudaSuccess) { if (multi) OPT_CUDA_CHECK(le); cudaGetLastError(); cgSeq = k0; return false; }
        return true;
    }
    size_t tileSmemFor(int a, int b, bool multi) const {
        const size_t nRing = (size_t)(b + 2) * (a + 2), nCap = (size_t)a * b;
        size_t bytes = ((nRing * 7 + nCap * 5) * sizeof(T) + nCap + 15) / 16 * 16;
        if (multi) bytes += (size_t)6 * a * sizeof(T);
        return (bytes + 15) / 16 * 16;
    }
    bool chooseTile(bool multi) {   // fills tileW/H/...; false when no shape fits
        int dev = 0, maxSmem = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int rows = D.y1 - D.y0, maxTiles = std::min(sms, kMaxLLCtas);
        auto fits = [&](int a, int b) {
            const long tiles = (long)((D.W + a - 1) / a) * ((rows + b - 1) / b);
            return tiles <= maxTiles && (long)a * b <= 4096 && tileSmemFor(a, b, multi) + 4096 <= (size_t)maxSmem;
        };
        int a = 0, b = 0;
        if (const char* spec = getenv("OPTB_RESIDENT_TILE")) { if (sscanf(spec, "%dx%d", &a, &b) != 2 || a < 1 || b < 1 || !fits(a, b)) return false; }
        else {
            if (!multi && !useCg(false)) return false;      // single GPU: the row-slab kernel is the default, tiles only on request
            bool found = false;
            const int widths[6] = { D.W, 1024, 512, 256, 128, 64 };
            for (int aw : widths) {   // widest tiles first: fewest ring cells per pixel for a given area
                if (aw > D.W) continue;
                for (int bh = 1; bh <= rows && !found; ++bh) if (fits(aw, bh)) { a = aw; b = bh; found = true; }
                if (found) break;
            }
            if (!found) return false;
        }
        tileW = a; tileH = b; tileAcross = (D.W + a - 1) / a; tileGrid = tileAcross * ((rows + b - 1) / b);
        tileSmem = tileSmemFor(a, b, multi); tileK = (int)(((size_t)a * b + 1023) / 1024); tileMulti = multi; tileChecked = nullptr;
        return true;
    }
    void* scratchMine = nullptr; void* scratchUp = nullptr; void* scratchDn = nullptr;
    void setPeerScratch(void* mine, void* up, void* dn) override { scratchMine = mine; scratchUp = up; scratchDn = dn; }
    bool pcgLoopMultiFits() override {   // asked once per connect, after setSlab(): tile shape + co-residency of all tiles
        if ((size_t)2 * 3 * D.W * LLWords<T>::N * sizeof(unsigned long long) > kPeerScratchBytes) return false;   // two LL rows must fit the peer scratch
        int dev = 0, coop = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (!coop || !chooseTile(true)) { tileW = -1; tileMulti = true; return false; }
        if (useCg(true)) { if (!cgPrepare(cgFnFor(true))) { tileW = -1; return false; } return true; }
        const void* fn = tileK <= 2 ? (const void*)k_iwPcgLoopTile<T, 2, true> : (const void*)k_iwPcgLoopTile<T, 4, true>;
        int occ = 0;
        if (tileK > 4 || cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tileSmem) != cudaSuccess) { cudaGetLastError(); tileW = -1; return false; }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 1024, tileSmem);
        if ((long)occ * sms < tileGrid) { tileW = -1; return false; }
        tileChecked = fn;
        return true;
    }
    bool launchTile(const LaunchCtx& c, int L, const HaloPlan* haloZ, T* delta, T* r, T* z, T* p, T* Ap, const T* pre) {
        const bool multi = haloZ != nullptr;
        if (tileW == 0 || tileMulti != multi) { if (!chooseTile(multi)) { tileW = -1; tileMulti = multi; } }
        if (tileW < 0) return false;
        IWTileMulti<T> M{};
        if (multi) {   // LL rows in the peer scratch areas: row 0 of a rank's scratch = data from its upper neighbour, row 1 = from its lower neighbour
            const size_t rowWords = (size_t)3 * D.W * LLWords<T>::N;
            unsigned long long* mine = (unsigned long long*)scratchMine;
            M.fromUp = haloZ->peerVec[0] ? mine : nullptr; M.fromDn = haloZ->peerVec[1] ? mine + rowWords : nullptr;
            M.toUp = haloZ->peerVec[0] ? (unsigned long long*)scratchUp + rowWords : nullptr;    // I am my upper neighbour's lower neighbour
            M.toDn = haloZ->peerVec[1] ? (unsigned long long*)scratchDn : nullptr;
            if ((haloZ->peerVec[0] && !scratchUp) || (haloZ->peerVec[1] && !scratchDn) || !mine) return false;
            if (useCg(true)) return launchCg(c, L, true, M, delta, r, z, p, Ap, pre);
            return tileK <= 2 ? launchTileK<2, true>(c, L, M, delta, r, z, p, Ap, pre) : launchTileK<4, true>(c, L, M, delta, r, z, p, Ap, pre);
        }
        if (useCg(false)) return launchCg(c, L, false, M, delta, r, z, p, Ap, pre);
        return tileK <= 2 ? launchTileK<2, false>(c, L, M, delta, r, z, p, Ap, pre) : launchTileK<4, false>(c, L, M, delta, r, z, p, Ap, pre);
    }
    bool launchResident(const LaunchCtx& c, int L, T* delta, T* r, T* z, T* p, T* Ap, const T* pre) {
        if (getenv("OPTB_NO_RESIDENT")) return false;
        if (launchTile(c, L, nullptr, delta, r, z, p, Ap, pre)) return true;
        if (resR == 0) {
            int dev = 0, sms = 148, maxSmem = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
            resR = (D.H + sms - 1) / sms; resGrid = (D.H + resR - 1) / resR;
            const size_t nRing = (size_t)(resR + 2) * D.W, nCap = (size_t)resR * D.W;
            resSmem = (nRing * 7 + nCap * 5) * sizeof(T) + nCap;
            resSmem = (resSmem + 15) / 16 * 16;
            resK = (int)((nCap + 1023) / 1024);
            if (resSmem > (size_t)maxSmem || resK > 4 || resGrid > kMaxLLCtas) resR = -1;
        }
        if (resR < 0) return false;
        return resK <= 2 ? launchResidentK<2>(c, L, delta, r, z, p, Ap, pre) : launchResidentK<4>(c, L, delta, r, z, p, Ap, pre);
    }
    // persistent whole-loop kernel (single GPU, GN): see k_iwPcgLoop
    LoopSync* loopSync = nullptr; unsigned long long loopSeq = 0; int gridLoop = 0; int coopOk = -1;
    bool pcgLoop(const LaunchCtx& c, int L, bool force, T* delta, T* r, T* z, T* p, T* p2, T* Ap, const T* pre, T** pFinal, const HaloPlan* haloZ) override {
        if (L <= 0) return false;
        if (haloZ) {   // multi-GPU: the plan only gets here when EVERY rank answered pcgLoopMultiFits() with yes, so the launch is collective
            if (!launchTile(c, L, haloZ, delta, r, z, p, Ap, pre)) { fprintf(stderr, "opt_b200: image_warping: the collective PCG-loop kernel could not be launched on this rank\n"); exit(4); }
            *pFinal = p;
            return true;
        }
        if (D.y0 != 0 || D.y1 != D.H) return false;
        if (coopOk < 0) {
            int dev = 0, coop = 0, sms = 148, occ = 0;
            cudaGetDevice(&dev); cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k_iwPcgLoop<T>, kBlock, 0);
            long g = (long)sms * occ; if (g > D.numTiles) g = D.numTiles; if (g > kMaxPartials) g = kMaxPartials;
            gridLoop = (int)g; coopOk = (coop && g >= 1) ? 1 : 0;
            if (coopOk) { OPT_CUDA_CHECK(cudaMalloc((void**)&loopSync, sizeof(LoopSync))); OPT_CUDA_CHECK(cudaMemset(loopSync, 0, sizeof(LoopSync))); }
        }
        if (!coopOk) return false;
        if (launchResident(c, L, delta, r, z, p, Ap, pre)) { *pFinal = p; return true; }   // state fits on chip: measured faster than the graph of per-step kernels
        if (!force) return false;   // the L2-state variant is slower than the graph: only on request
        unsigned long long k0 = loopSeq; loopSeq += 2ull * (unsigned long long)L;
        ReduceWorkspace ws = c.ws;
        void* args[] = { (void*)&D, (void*)&ws, (void*)&loopSync, (void*)&k0, (void*)&L, (void*)&delta, (void*)&r, (void*)&z, (void*)&p, (void*)&p2, (void*)&Ap, (void*)&pre };
        OPT_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_iwPcgLoop<T>, dim3((unsigned)gridLoop), dim3(kBlock), args, 0, c.stream));
        *pFinal = ((L - 1) & 1) ? p2 : p;
        return true;
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan* halo) override {
        if constexpr (std::is_same<T, float>::value) {
            if (tmaOk) {
                const IWMaps& m = mapsFor(x, nullptr);
                const HaloPlan* hp = halo;
                if (upperRows > 0 || lowerRows > 0) {
                    if (CtC) k_iwApplyJTJTma<false, true, kStagesMulti><<<gridTma[0], kBlock, kStagesMulti * kStageBytesPlain, c.stream>>>(D, m, c.ws, dotSlot, 0, 0, nullptr, out, CtC, tmaTilesX, tmaTiles, x, hp);
                    else launchPdl(k_iwApplyJTJTma<false, false, kStagesMulti>, gridTma[0], kBlock, (size_t)(kStagesMulti * kStageBytesPlain), c.stream, c.pdl, D, m, c.ws, dotSlot, 0, 0, (float*)nullptr, out, (const float*)nullptr, tmaTilesX, tmaTiles, (const float*)x, hp);
                } else {
                    if (CtC) k_iwApplyJTJTma<false, true, kStagesSingle><<<gridTma[1], kBlock, kStagesSingle * kStageBytesPlain, c.stream>>>(D, m, c.ws, dotSlot, 0, 0, nullptr, out, CtC, tmaTilesX, tmaTiles, x, hp);
                    else launchPdl(k_iwApplyJTJTma<false, false, kStagesSingle>, gridTma[1], kBlock, (size_t)(kStagesSingle * kStageBytesPlain), c.stream, c.pdl, D, m, c.ws, dotSlot, 0, 0, (float*)nullptr, out, (const float*)nullptr, tmaTilesX, tmaTiles, (const float*)x, hp);
                }
                return;
            }
        }
        if (CtC) k_iwApplyJTJ<T, false, true><<<gridJTJ, kBlock, 0, c.stream>>>(D, c.ws, dotSlot, 0, 0, x, nullptr, nullptr, out, CtC);
        else k_iwApplyJTJ<T, false, false><<<gridJTJ, kBlock, 0, c.stream>>>(D, c.ws, dotSlot, 0, 0, x, nullptr, nullptr, out, nullptr);
    }
    bool step3ApplyJTJ(const LaunchCtx& c, const T* z, const T* pOld, T* pNew, T* out, const T* CtC, int rzNew, int rzOld, int dotSlot, const HaloPlan* halo) override {
        if (dotSlot == -2) return true;
        if constexpr (std::is_same<T, float>::value) {
            if (tmaOk) {
                const IWMaps& m = mapsFor(pOld, z);
                const HaloPlan* hp = halo;
                if (upperRows > 0 || lowerRows > 0) {
                    if (CtC) k_iwApplyJTJTma<true, true, kStagesMulti><<<gridTmaF[0], kBlock, kStagesMulti * kStageBytesFused, c.stream>>>(D, m, c.ws, dotSlot, rzNew, rzOld, pNew, out, CtC, tmaTilesX, tmaTiles, z, hp);
                    else launchPdl(k_iwApplyJTJTma<true, false, kStagesMulti>, gridTmaF[0], kBlock, (size_t)(kStagesMulti * kStageBytesFused), c.stream, c.pdl, D, m, c.ws, dotSlot, rzNew, rzOld, pNew, out, (const float*)nullptr, tmaTilesX, tmaTiles, (const float*)z, hp);
                } else {
                    if (CtC) k_iwApplyJTJTma<true, true, kStagesSingle><<<gridTmaF[1], kBlock, kStagesSingle * kStageBytesFused, c.stream>>>(D, m, c.ws, dotSlot, rzNew, rzOld, pNew, out, CtC, tmaTilesX, tmaTiles, z, hp);
                    else launchPdl(k_iwApplyJTJTma<true, false, kStagesSingle>, gridTmaF[1], kBlock, (size_t)(kStagesSingle * kStageBytesFused), c.stream, c.pdl, D, m, c.ws, dotSlot, rzNew, rzOld, pNew, out, (const float*)nullptr, tmaTilesX, tmaTiles, (const float*)z, hp);
                }
                return true;
            }
        }
        if (CtC) k_iwApplyJTJ<T, true, true><<<gridJTJF, kBlock, 0, c.stream>>>(D, c.ws, dotSlot, rzNew, rzOld, pOld, z, pNew, out, CtC);
        else k_iwApplyJTJ<T, true, false><<<gridJTJF, kBlock, 0, c.stream>>>(D, c.ws, dotSlot, rzNew, rzOld, pOld, z, pNew, out, nullptr);
        return true;
    }
    void cost(const LaunchCtx& c) override { k_iwCost<T, false><<<gridCost, kBlock, 0, c.stream>>>(D, c.ws, kCost, nullptr); }
    void modelCost(const LaunchCtx& c, const T* delta) override { k_iwCost<T, true><<<gridCost, kBlock, 0, c.stream>>>(D, c.ws, kModelCost, delta); }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override { k_iwCtC<T><<<gridJTJ, kBlock, 0, c.stream>>>(D, CtC, invRadius); }
};

template <class T> EnergyOps<T>* makeImageWarping(const unsigned* dims) { return new ImageWarpingOps<T>(dims); }
template EnergyOps<float>* makeImageWarping<float>(const unsigned*);
template EnergyOps<double>* makeImageWarping<double>(const unsigned*);

}  // namespace optb
