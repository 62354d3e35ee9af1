// sfs_tile.cuh -- tile-level building blocks of the shape_from_shading kernel set: the operators of sfs_device.cuh evaluated for a
// tw x th tile of pixels out of staged planes instead of pixel by pixel out of global memory.
//
// Every operator of this energy is "one value per ROW (y = row . vec, or the row's residual F), then for every unknown u the sum over
// the rows that read X(u) of coefficient x value" (sfsGather in sfs_device.cuh).  The per-pixel form re-derives each row value once per
// unknown that row touches (5x for r_gR / r_gC, 5x3 for r_s); the tile form computes it once per row centre in (tile + 1-pixel ring):
//     sp   [(th+4) x (tw+4)]   the vector the rows are applied to, tile + 2-pixel ring (0 outside the image)
//     sc   [10][cells]         the r_gR / r_gC coefficients of the row centred at the cell          cells = (th+2) x (tw+2)
//     sy   [5][cells]          the five row values (r_gR, r_gC, r_s x/y/z) of the row centred at the cell
//     sfl  [cells]             flag byte of the cell (0 outside the image)
//     sAx, sAy                 back-projection factors (x - u_x)/f_x, (y - u_y)/f_y of the columns / rows of tile + 2-pixel ring
//     sd   [(th+4) x (tw+4)]   (whole-loop kernel only) a second vector plane: delta, for LM's residual-reset iterations
// Used by k_sfsInitTiled (PCGInit1 + PCGComputeCtC + PCGFinalizeDiagonal in one pass) and by the whole-loop kernel k_sfsLoopResident
// (energy_sfs.cu); the functions are plain host/device code so tests/host/sfs_tile_host.cpp can run both algorithms on the CPU, "thread"
// by "thread", against the oracle (tests/test_sfs_host_parity.py).  Same products in the same order as sfsGather<T, MODE>.
#pragma once
#include <cstddef>
#if !defined(__CUDACC__)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#endif
#include "sfs_device.cuh"

namespace optb {

template <class T>
struct SfsTileView {
    int x0, y0, tw, th;          // image coordinates of the tile's first pixel, nominal extent
    T* sp; T* sc; T* sy; unsigned char* sfl; T* sAx; T* sAy;
    T* sd;                       // whole-loop kernel only: delta with a 2-pixel ring, for the residual-reset iterations (nullptr otherwise)
    unsigned mTW, mRW, mPW;      // ceil(2^32 / d) for d = tw, tw + 2, tw + 4: index -> (row, column) without an integer division (exact below 2^16)
    SFS_HD int RW() const { return tw + 2; }
    SFS_HD int RH() const { return th + 2; }
    SFS_HD int PW() const { return tw + 4; }
    SFS_HD int PH() const { return th + 4; }
    SFS_HD int cells() const { return (tw + 2) * (th + 2); }
};
SFS_HD unsigned sfsMagic(int d) { return (unsigned)((0x100000000ull + (unsigned long long)d - 1ull) / (unsigned long long)d); }
SFS_HD int sfsFastDiv(int c, unsigned magic) { return (int)(((unsigned long long)(unsigned)c * magic) >> 32); }   // c / d for 0 <= c < 2^16, 2 <= d < 2^16
// bytes of the planes above for a tw x th tile (sp, sc, sy, the two factor tables, then the flag bytes), each plane 16-byte aligned
SFS_HD constexpr size_t sfsUp16(size_t v) { return (v + 15) / 16 * 16; }
template <class T> SFS_HD constexpr size_t sfsTilePlaneBytes(int tw, int th, bool withDelta = false) {
    return sfsUp16((size_t)(tw + 4) * (th + 4) * sizeof(T)) * (withDelta ? 2 : 1) + sfsUp16(10 * (size_t)(tw + 2) * (th + 2) * sizeof(T))
         + sfsUp16(5 * (size_t)(tw + 2) * (th + 2) * sizeof(T)) + sfsUp16((size_t)(tw + 4) * sizeof(T)) + sfsUp16((size_t)(th + 4) * sizeof(T)) + sfsUp16((size_t)(tw + 2) * (th + 2));
}
template <class T> SFS_HD SfsTileView<T> sfsCarveTile(unsigned char* base, int x0, int y0, int tw, int th, bool withDelta = false) {
    SfsTileView<T> V; V.x0 = x0; V.y0 = y0; V.tw = tw; V.th = th;
    V.mTW = sfsMagic(tw); V.mRW = sfsMagic(tw + 2); V.mPW = sfsMagic(tw + 4);
    const size_t cells = (size_t)(tw + 2) * (th + 2), pcells = (size_t)(tw + 4) * (th + 4);
    V.sp = reinterpret_cast<T*>(base); base += sfsUp16(pcells * sizeof(T));
    V.sd = nullptr;
    if (withDelta) { V.sd = reinterpret_cast<T*>(base); base += sfsUp16(pcells * sizeof(T)); }
    V.sc = reinterpret_cast<T*>(base); base += sfsUp16(10 * cells * sizeof(T));
    V.sy = reinterpret_cast<T*>(base); base += sfsUp16(5 * cells * sizeof(T));
    V.sAx = reinterpret_cast<T*>(base); base += sfsUp16((size_t)(tw + 4) * sizeof(T));
    V.sAy = reinterpret_cast<T*>(base); base += sfsUp16((size_t)(th + 4) * sizeof(T));
    V.sfl = base;
    return V;
}

// ---- staging (once per linearisation): factor tables, then flag byte + the ten coefficients of row cell c
template <class T> SFS_HD void sfsTileFactors(const SfsData<T>& P, const SfsTileView<T>& V, int i) {   // i in [0, PW + PH)
    if (i < V.PW()) V.sAx[i] = sfsAx(P, V.x0 - 2 + i);
    else if (i < V.PW() + V.PH()) V.sAy[i - V.PW()] = sfsAy(P, V.y0 - 2 + (i - V.PW()));
}
template <class T> SFS_HD void sfsTileStageCell(const SfsData<T>& P, const SfsTileView<T>& V, int c) {
    const int RW = V.RW(), cells = V.cells();
    const int ly = sfsFastDiv(c, V.mRW), lx = c - ly * RW, gx = V.x0 - 1 + lx, gy = V.y0 - 1 + ly;
    const long N = sfsN(P);
    unsigned f = 0;
    T co[10] = { T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0) };
    if (sfsInImage(P, gx, gy)) {
        const long ci = (long)gy * P.W + gx;
        f = P.fl[ci];
        if (f & kSfsInner) {
            SFS_UNROLL for (int m = 0; m < 5; ++m) { co[m] = P.cR[(long)m * N + ci]; co[5 + m] = P.cC[(long)m * N + ci]; }
        }
    }
    V.sfl[c] = (unsigned char)f;
    SFS_UNROLL for (int k = 0; k < 10; ++k) V.sc[k * cells + c] = co[k];
}

// ---- row values of cell c:  y = row . sp   (rows with a zero coefficient contribute an exact 0 instead of being skipped)
template <class T> SFS_HD void sfsTileRowsTimesVec(const SfsData<T>& P, const SfsTileView<T>& V, int c, const T* vec) {   // vec: V.sp or V.sd
    const int RW = V.RW(), PW = V.PW(), cells = V.cells();
    const int ly = sfsFastDiv(c, V.mRW), lx = c - ly * RW;
    const unsigned f = V.sfl[c];
    T y[5] = { T(0), T(0), T(0), T(0), T(0) };
    if (f & (kSfsInner | kSfsFlat)) {
        // the seven entries the three row types read between them (dependency tables of sfs_device.cuh), loaded once
        const T* v0 = vec + (ly + 1) * PW + lx + 1;
        const T c00 = v0[0], cm0 = v0[-1], cp0 = v0[1], c0m = v0[-PW], c0p = v0[PW], cpm = v0[1 - PW], cmp = v0[PW - 1];
        if (f & kSfsInner) {
            const T* a = V.sc + c;   // r_gR: (0,0) (-1,0) (0,-1) (1,0) (1,-1)     r_gC: (0,0) (-1,0) (0,-1) (0,1) (-1,1)
            y[0] += a[0] * c00; y[0] += a[cells] * cm0; y[0] += a[2 * cells] * c0m; y[0] += a[3 * cells] * cp0; y[0] += a[4 * cells] * cpm;
            y[1] += a[5 * cells] * c00; y[1] += a[6 * cells] * cm0; y[1] += a[7 * cells] * c0m; y[1] += a[8 * cells] * c0p; y[1] += a[9 * cells] * cmp;
        }
        if (f & kSfsFlat) {                                                        // r_s: (0,0) (-1,0) (0,-1) (1,0) (0,1)
            // row k of r_s is ws (4 P_k(c) - P_k(left) - P_k(up) - P_k(right) - P_k(down)) with P = (a_x d, a_y d, d): the factor of a column
            // (row) is shared by the pixels above / below (left / right) of the centre, so the 15 products regroup into 7 + 7 + 5 operations
            const T ax0 = V.sAx[lx], ax1 = V.sAx[lx + 1], ax2 = V.sAx[lx + 2], ay0 = V.sAy[ly], ay1 = V.sAy[ly + 1], ay2 = V.sAy[ly + 2];   // columns gx-1..gx+1, rows gy-1..gy+1
            const T c4 = T(4) * c00;
            y[2] = P.ws * (ax1 * (c4 - c0m - c0p) - ax0 * cm0 - ax2 * cp0);
            y[3] = P.ws * (ay1 * (c4 - cm0 - cp0) - ay0 * c0m - ay2 * c0p);
            y[4] = P.ws * (c4 - cm0 - c0m - cp0 - c0p);
        }
    }
    SFS_UNROLL for (int k = 0; k < 5; ++k) V.sy[k * cells + c] = y[k];
}
// ---- row values of cell c:  y = F(row), the residuals at the current X (sfsFR, sfsFC, sfsFS of sfs_device.cuh, factors from the tables)
template <class T> SFS_HD void sfsTileRowResiduals(const SfsData<T>& P, const SfsTileView<T>& V, int c) {
    const int RW = V.RW(), cells = V.cells();
    const int ly = sfsFastDiv(c, V.mRW), lx = c - ly * RW, gx = V.x0 - 1 + lx, gy = V.y0 - 1 + ly;
    const unsigned f = V.sfl[c];
    T y[5] = { T(0), T(0), T(0), T(0), T(0) };
    if (f & (kSfsInner | kSfsFlat)) {   // either flag implies an in-image cell
        const long ci = (long)gy * P.W + gx;
        if (f & kSfsInner) {
            const T Bc = P.B[ci];
            y[0] = P.wg * ((Bc - P.B[ci + 1]) * T(P.eR[ci]));
            y[1] = P.wg * ((Bc - P.B[ci + P.W]) * T(P.eC[ci]));
        }
        if (f & kSfsFlat) {
            const T ax[3] = { V.sAx[lx], V.sAx[lx + 1], V.sAx[lx + 2] }, ay[3] = { V.sAy[ly], V.sAy[ly + 1], V.sAy[ly + 2] };
            SFS_UNROLL for (int m = 0; m < 5; ++m) {
                const T xm = P.X[ci + (long)sfsDy(2, m) * P.W + sfsDx(2, m)];
                y[2] += sfsCoefFromA(P, ax[1 + sfsDx(2, m)], m) * xm;
                y[3] += sfsCoefFromA(P, ay[1 + sfsDy(2, m)], m) * xm;
                y[4] += sfsCoefFromA(P, T(1), m) * xm;
            }
        }
    }
    SFS_UNROLL for (int k = 0; k < 5; ++k) V.sy[k * cells + c] = y[k];
}

// ---- gathers for the tile pixel (ty, tx) (an INCLUDED pixel; the caller checks).  `first` is the r_p term:
//      wp (wp x_u) for J^T J x, wp F_p(u) for J^T F.
template <class T> SFS_HD T sfsTileGather(const SfsData<T>& P, const SfsTileView<T>& V, int ty, int tx, T first) {
    const int RW = V.RW(), cells = V.cells();
    T o = first;
    SFS_UNROLL for (int t = 0; t < 2; ++t)
        SFS_UNROLL for (int k = 0; k < 5; ++k) {                                  // u is dependency k of the row centred at u - offset_k
            const int cc = (ty + 1 - sfsDy(t, k)) * RW + tx + 1 - sfsDx(t, k);
            o += V.sc[(t * 5 + k) * cells + cc] * V.sy[t * cells + cc];
        }
    // r_s: dependency k of the row centred at u - offset_k is u itself, so all fifteen coefficients carry u's own factors:
    // sum_k coef_k y_k = ws a_u (4 y(u) - y(u+left..) ...) -- one 5-point combination per component, then the three factors
    const T axu = V.sAx[tx + 2], ayu = V.sAy[ty + 2];
    const int c0 = (ty + 1) * RW + tx + 1;
    const T* s2 = V.sy + 2 * cells + c0; const T* s3 = V.sy + 3 * cells + c0; const T* s4 = V.sy + 4 * cells + c0;
    // offsets of the five centres u - offset_k, k = 0..4: (0,0) (+1,0) (0,+1) (-1,0) (0,-1)
    const T l2 = T(4) * s2[0] - s2[1] - s2[RW] - s2[-1] - s2[-RW];
    const T l3 = T(4) * s3[0] - s3[1] - s3[RW] - s3[-1] - s3[-RW];
    const T l4 = T(4) * s4[0] - s4[1] - s4[RW] - s4[-1] - s4[-RW];
    o += P.ws * (axu * l2 + ayu * l3 + l4);
    return o;
}
template <class T> SFS_HD T sfsTileGatherDiag(const SfsData<T>& P, const SfsTileView<T>& V, int ty, int tx) {   // diag J^T J, order of sfsGather<T, 2>
    const int RW = V.RW(), cells = V.cells();
    T acc = P.wp * P.wp;
    SFS_UNROLL for (int t = 0; t < 2; ++t)
        SFS_UNROLL for (int k = 0; k < 5; ++k) {
            const T a = V.sc[(t * 5 + k) * cells + (ty + 1 - sfsDy(t, k)) * RW + tx + 1 - sfsDx(t, k)];
            acc += a * a;
        }
    const T aU[3] = { V.sAx[tx + 2], V.sAy[ty + 2], T(1) };
    SFS_UNROLL for (int k = 0; k < 5; ++k) {
        if (!(V.sfl[(ty + 1 - sfsDy(2, k)) * RW + tx + 1 - sfsDx(2, k)] & kSfsFlat)) continue;
        SFS_UNROLL for (int comp = 0; comp < 3; ++comp) { const T a = sfsCoefFromA(P, aU[comp], k); acc += a * a; }
    }
    return acc;
}

// =============================================================================================================================
// precompute (B, flags, the r_gR / r_gC row coefficients) AND computeCost for a tw x th tile in one pass (k_sfsPrecomputeTiled): B and its
// partials are formed for the tile plus the column to its right and the row below it (what the rows centred in the tile read; the
// neighbouring tiles form the same values from the same inputs for their own pixels) and kept in shared memory:
//     sB [4][(th+1) x (tw+1)]   B, dB/dX(c), dB/dX(c-(1,0)), dB/dX(c-(0,1))        sF [(th+1) x (tw+1)]   flag bytes (tile pixels only)
// The per-pixel path does this in two launches with dB going through global memory (k_sfsPrecomputeB, k_sfsBuildRows) plus k_sfsCost.
template <class T>
struct SfsPreView {
    int x0, y0, tw, th; T* sB; unsigned char* sF; unsigned mBW;
    SFS_HD int BW() const { return tw + 1; }
    SFS_HD int bcells() const { return (tw + 1) * (th + 1); }
};
template <class T> SFS_HD constexpr size_t sfsPrePlaneBytes(int tw, int th) { return sfsUp16(4 * (size_t)(tw + 1) * (th + 1) * sizeof(T)) + sfsUp16((size_t)(tw + 1) * (th + 1)); }
template <class T> SFS_HD SfsPreView<T> sfsCarvePre(unsigned char* base, int x0, int y0, int tw, int th) {
    SfsPreView<T> V; V.x0 = x0; V.y0 = y0; V.tw = tw; V.th = th; V.mBW = sfsMagic(tw + 1);
    V.sB = reinterpret_cast<T*>(base); V.sF = base + sfsUp16(4 * (size_t)(tw + 1) * (th + 1) * sizeof(T));
    return V;
}
// pass 1, cell c of the (th+1) x (tw+1) region: B and partials into the planes; tile pixels also get B and the flag byte in global memory
template <class T> SFS_HD void sfsPreCell(const SfsData<T>& P, const SfsPreView<T>& V, int c) {
    const int BW = V.BW(), n = V.bcells();
    const int ly = sfsFastDiv(c, V.mBW), lx = c - ly * BW, gx = V.x0 + lx, gy = V.y0 + ly;
    const SfsDual<T> b = sfsB(P, gx, gy);                       // 0 outside the image
    V.sB[c] = b.v; V.sB[n + c] = b.d0; V.sB[2 * n + c] = b.d1; V.sB[3 * n + c] = b.d2;
    unsigned f = 0;
    if (lx < V.tw && ly < V.th && sfsInImage(P, gx, gy)) {
        const long ci = (long)gy * P.W + gx;
        f = sfsFlagsAt(P, gx, gy);
        P.B[ci] = b.v; P.fl[ci] = (unsigned char)f;
    }
    V.sF[c] = (unsigned char)f;
}
// pass 2, tile pixel (ty, tx) inside the image: the ten row coefficients (sfsBuildRows) and 1/2 sum of squares of the six rows (sfsCostAt)
template <class T> SFS_HD T sfsPrePixel(const SfsData<T>& P, const SfsPreView<T>& V, int ty, int tx) {
    const int BW = V.BW(), n = V.bcells();
    const int c0 = ty * BW + tx, cr = c0 + 1, cd = c0 + BW;
    const int gx = V.x0 + tx, gy = V.y0 + ty;
    const long ci = (long)gy * P.W + gx, N = sfsN(P);
    const unsigned f = V.sF[c0];
    T r[5] = { T(0), T(0), T(0), T(0), T(0) }, q[5] = { T(0), T(0), T(0), T(0), T(0) };
    T e = T(0);
    if (f & kSfsIncluded) { const T fp = P.wp * (P.X[ci] - P.D[ci]); e += fp * fp; }
    if (f & kSfsInner) {
        const T d0 = V.sB[n + c0], d1 = V.sB[2 * n + c0], d2 = V.sB[3 * n + c0];
        const T eR = T(P.eR[ci]), eC = T(P.eC[ci]);
        {   // B(c) - B(c+(1,0)):  B(c+(1,0)) reads X(c+(1,0)) [slot 0], X(c) [slot 1], X(c+(1,-1)) [slot 2]
            const T s = P.wg * eR;
            r[0] = s * (d0 - V.sB[2 * n + cr]); r[1] = s * d1; r[2] = s * d2; r[3] = -s * V.sB[n + cr]; r[4] = -s * V.sB[3 * n + cr];
        }
        {   // B(c) - B(c+(0,1)):  B(c+(0,1)) reads X(c+(0,1)) [slot 0], X(c+(-1,1)) [slot 1], X(c) [slot 2]
            const T s = P.wg * eC;
            q[0] = s * (d0 - V.sB[3 * n + cd]); q[1] = s * d1; q[2] = s * d2; q[3] = -s * V.sB[n + cd]; q[4] = -s * V.sB[2 * n + cd];
        }
        if (f & kSfsIncluded) {
            const T fr = P.wg * ((V.sB[c0] - V.sB[cr]) * eR), fc = P.wg * ((V.sB[c0] - V.sB[cd]) * eC);
            e += fr * fr; e += fc * fc;
        }
    }
    if ((f & kSfsFlat) && (f & kSfsIncluded)) {
        SFS_UNROLL for (int comp = 0; comp < 3; ++comp) { const T fs = sfsFSWithFlag(P, gx, gy, comp); e += fs * fs; }
    }
    SFS_UNROLL for (int k = 0; k < 5; ++k) { P.cR[(long)k * N + ci] = r[k]; P.cC[(long)k * N + ci] = q[k]; }
    return T(0.5) * e;
}

// =============================================================================================================================
// PCGInit1 for one pixel out of the gathers above (+ PCGComputeCtC + PCGFinalizeDiagonal when LM): what k_sfsInit, k_sfsCtC and
// k_lmFinalizeDiagonal (solver.cu) do in three passes over memory.  Returns the pixel's contribution to r . p.
template <class T> SFS_HD T sfsGuardedInvertT(T v) { const T s = T(1) + sfsSqrt(v); return T(1) / (s * s); }   // CERES variant, solverGPUGaussNewton.t:323-332
template <class T>
struct SfsInitOut { T* delta; T* r; T* pre; T* p; T* CtC; T* SSq; T* b; };
template <class T, bool LM>
SFS_HD T sfsInitPixel(bool included, T jtf, T diag, long i, const SfsInitOut<T>& O, int usePre, int saveSSq, T radius, T minDiag, T maxDiag) {
    T rr = T(0), pr = T(0);
    if (included) { rr = -jtf; pr = sfsGuardedInvertT(usePre ? diag : T(1)); }   // PCGInit1 order (:379-385): without a preconditioner the constant 1 is inverted too
    O.delta[i] = T(0); O.r[i] = rr;
    if (!LM) { const T pp = pr * rr; O.pre[i] = pr; O.p[i] = pp; return rr * pp; }
    // PCGSaveSSq (first nonlinear iteration, :1038-1041), then PCGFinalizeDiagonal (:631-664) as k_lmFinalizeDiagonal applies it
    T ssq = saveSSq ? pr : O.SSq[i];
    if (saveSSq) O.SSq[i] = pr;
    else if (pr > T(0) && !(ssq > T(0))) { ssq = pr; O.SSq[i] = ssq; }
    T acc = T(0);
    if (pr > T(0) && ssq > T(0)) {
        const T invRadius = T(1) / radius;
        const T unclamped = (included ? diag : T(0)) * invRadius;
        const T invS = T(1) / ssq;
        const T clampMul = invS / radius;
        const T mn = minDiag * clampMul, mx = maxDiag * clampMul;
        const T c = unclamped < mn ? mn : (unclamped > mx ? mx : unclamped);
        O.CtC[i] = c;
        const T pr2 = T(1) / (c + radius * unclamped);
        O.pre[i] = pr2; O.b[i] = rr;
        const T pp = pr2 * rr;
        O.p[i] = pp;
        acc = rr * pp;
    } else {
        O.CtC[i] = T(0); O.pre[i] = T(0); O.b[i] = T(0); O.p[i] = T(0);
    }
    return acc;
}

// =============================================================================================================================
// Whole PCG loop with the state of a tile on chip (k_sfsLoopResident).  One CTA per tile: the planes of SfsTileView hold the
// direction p (tile + 2-pixel ring), the coefficients and the flags; delta, r, b, CtC, pre, z and Ap of a thread's (at most KMAX)
// pixels live in registers.  Per PCG step only the ring of z crosses CTAs (through global memory / L2) plus two all-reduces.
//   pixel q of thread tid:  c = tid + q * NT,  (ty, tx) = (c / tw, c % tw),  valid if c < tw * th and inside the image.
template <class T, int KMAX>
struct SfsLoopRegs { T delta[KMAX], r[KMAX], b[KMAX], ctc[KMAX], pre[KMAX], z[KMAX], ap[KMAX]; };

template <class T> SFS_HD bool sfsLoopPixel(const SfsData<T>& P, const SfsTileView<T>& V, int c, int& ty, int& tx, long& gi) {
    if (c >= V.tw * V.th) return false;
    ty = sfsFastDiv(c, V.mTW); tx = c - ty * V.tw;
    const int gx = V.x0 + tx, gy = V.y0 + ty;
    if (gx >= P.W || gy >= P.H) return false;
    gi = (long)gy * P.W + gx;
    return true;
}
// Tile-edge values cross CTAs as self-validating words (payload | 32-bit tag in one 8-byte store, llStore / llLoad of common.cuh): the
// reader polls the word until it carries the tag of the exchange it is in, so no fence and no barrier orders the data -- the two
// all-reduces of a PCG step stay pure reductions.  One word (float) or two (double) per pixel, indexed like the image.
// Tags grow monotonically per slot: 8 x (all-reduces completed so far) + {1: z of this step, 2: delta of a reset step, 3: delta for the model
// cost, 4: the first direction when the iteration's head runs inside the kernel}.
#if defined(__CUDA_ARCH__)
template <class T> SFS_HD void sfsEdgeStore(unsigned long long* ll, long gi, T v, unsigned tag) { llStore(ll + gi * LLWords<T>::N, v, tag); }
template <class T> SFS_HD T sfsEdgeLoad(const unsigned long long* ll, long gi, unsigned tag) { T v; llLoad(ll + gi * LLWords<T>::N, tag, v); return v; }
#else
// host emulation (tests/host/sfs_tile_host.cpp; two words per pixel: value, tag): phases run in lock step, so a word whose tag does not
// match is a logic error
template <class T> inline void sfsEdgeStore(unsigned long long* ll, long gi, T v, unsigned tag) {
    double d = (double)v; unsigned long long bits; memcpy(&bits, &d, 8);
    ll[gi * 2] = bits; ll[gi * 2 + 1] = tag;
}
template <class T> inline T sfsEdgeLoad(const unsigned long long* ll, long gi, unsigned tag) {
    if ((unsigned)ll[gi * 2 + 1] != tag || tag == 0u) { fprintf(stderr, "edge word %ld carries tag %u, expected %u\n", gi, (unsigned)ll[gi * 2 + 1], tag); abort(); }
    double d; memcpy(&d, &ll[gi * 2], 8); return (T)d;
}
#endif
// kinds 1..4, never 0: the buffers start zeroed, and a reader of the plan's very first exchange (no all-reduce done yet) must not take an
// untouched word for a published one (it did with tag 4 k + 0: the first nonlinear iteration of a plan read zeros for some ring cells)
SFS_HD unsigned sfsEdgeTag(unsigned long long reductionsDone, int what) { return (unsigned)(8ull * reductionsDone + (unsigned long long)what); }
constexpr int kSfsEdgeZ = 1, kSfsEdgeDeltaReset = 2, kSfsEdgeDeltaModel = 3, kSfsEdgeFirstDirection = 4;

// prologue: the direction with its ring, this thread's pixels
template <class T, int KMAX>
SFS_HD void sfsLoopLoad(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT, const T* delta, const T* r, const T* b,
                        const T* CtC, const T* pre, const T* p, int usePre, int lm) {
    const int PW = V.PW(), pcells = PW * V.PH();
    for (int c = tid; c < pcells; c += NT) {
        const int ly = sfsFastDiv(c, V.mPW), lx = c - ly * PW, gx = V.x0 - 2 + lx, gy = V.y0 - 2 + ly;
        V.sp[c] = sfsInImage(P, gx, gy) ? p[(long)gy * P.W + gx] : T(0);
        if (V.sd) V.sd[c] = T(0);
    }
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        R.delta[q] = R.r[q] = R.b[q] = R.ctc[q] = R.z[q] = R.ap[q] = T(0); R.pre[q] = T(1);
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            R.delta[q] = delta[gi]; R.r[q] = r[gi];
            if (lm) { R.b[q] = b[gi]; R.ctc[q] = CtC[gi]; }
            if (usePre) R.pre[q] = pre[gi];
        }
    }
}
// head of the nonlinear iteration INSIDE the loop kernel (LM): PCGInit1 + PCGSaveSSq / PCGComputeCtC / PCGFinalizeDiagonal for the thread's
// pixels, straight into the loop registers (what k_sfsInitTiled writes to memory and sfsLoopLoad reads back).  The row planes must hold
// the residuals (sfsTileRowResiduals) and the vector planes zeros.  The initial direction goes into the tile's plane and, for pixels
// within two of the border, into the edge words (tag kind 0) for the neighbours' rings.  Returns the thread's part of r . p.
template <class T, int KMAX>
SFS_HD T sfsLoopHead(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT, int usePre, int saveSSq, T radius, T minDiag, T maxDiag,
                     T* SSq, unsigned long long* zLL, unsigned tag) {
    const int PW = V.PW(), RW = V.RW();
    T acc = T(0);
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        R.delta[q] = R.r[q] = R.b[q] = R.ctc[q] = R.z[q] = R.ap[q] = T(0); R.pre[q] = T(0);
        if (!sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) continue;
        const bool inc = (V.sfl[(ty + 1) * RW + tx + 1] & kSfsIncluded) != 0;
        T rr = T(0), pr = T(0), dg = T(0);
        if (inc) {
            rr = -sfsTileGather(P, V, ty, tx, P.wp * sfsFP(P, gi));
            dg = sfsTileGatherDiag(P, V, ty, tx);
            pr = sfsGuardedInvertT(usePre ? dg : T(1));
        }
        T ssq = saveSSq ? pr : SSq[gi];
        if (saveSSq) SSq[gi] = pr;
        else if (pr > T(0) && !(ssq > T(0))) { ssq = pr; SSq[gi] = ssq; }
        T pp = T(0);
        if (pr > T(0) && ssq > T(0)) {
            const T invRadius = T(1) / radius;
            const T unclamped = dg * invRadius;
            const T invS = T(1) / ssq;
            const T clampMul = invS / radius;
            const T mn = minDiag * clampMul, mx = maxDiag * clampMul;
            const T c = unclamped < mn ? mn : (unclamped > mx ? mx : unclamped);
            const T pr2 = T(1) / (c + radius * unclamped);
            pp = pr2 * rr;
            R.ctc[q] = c; R.pre[q] = pr2; R.b[q] = rr;
            acc += rr * pp;
        }
        R.r[q] = rr;
        V.sp[(ty + 2) * PW + tx + 2] = pp;
        if (ty < 2 || ty >= V.th - 2 || tx < 2 || tx >= V.tw - 2) sfsEdgeStore<T>(zLL, gi, pp, tag);
    }
    return acc;
}
// PCGStep3 (:537-550) for the thread's own pixels (z in registers) and, for the ring cells, from the z their owners published
template <class T, int KMAX>
SFS_HD void sfsLoopUpdateDirection(const SfsData<T>& P, const SfsTileView<T>& V, const SfsLoopRegs<T, KMAX>& R, int tid, int NT, T beta, const unsigned long long* zLL, unsigned tag) {
    const int PW = V.PW(), PH = V.PH();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) { const int cc = (ty + 2) * PW + tx + 2; V.sp[cc] = R.z[q] + beta * V.sp[cc]; }
    }
    // ring: two rows above, two below, two columns left and right of the tile
    const int nRing = 4 * PW + 4 * V.th;
    for (int j = tid; j < nRing; j += NT) {
        int ly, lx;
        if (j < 2 * PW) { ly = j >= PW ? 1 : 0; lx = j - ly * PW; }
        else if (j < 4 * PW) { const int jj = j - 2 * PW; ly = jj >= PW ? 1 : 0; lx = jj - ly * PW; ly += PH - 2; }
        else { const int jj = j - 4 * PW; ly = 2 + (jj >> 2); const int s = jj & 3; lx = s < 2 ? s : PW - 4 + s; }
        const int gx = V.x0 - 2 + lx, gy = V.y0 - 2 + ly;
        if (sfsInImage(P, gx, gy)) { const int cc = ly * PW + lx; V.sp[cc] = sfsEdgeLoad<T>(zLL, (long)gy * P.W + gx, tag) + beta * V.sp[cc]; }
    }
}
// PCGStep1 (:421-434), second half: Ap for the thread's pixels out of the row values; returns the thread's part of p . Ap
template <class T, int KMAX>
SFS_HD T sfsLoopApply(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT) {
    const int PW = V.PW(), RW = V.RW();
    T acc = T(0);
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        T o = T(0);
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi) && (V.sfl[(ty + 1) * RW + tx + 1] & kSfsIncluded)) {
            const T xi = V.sp[(ty + 2) * PW + tx + 2];
            o = sfsTileGather(P, V, ty, tx, P.wp * (P.wp * xi));
            o += R.ctc[q] * xi;                                   // LM: + CtC p (0 for Gauss-Newton)
            acc += xi * o;
        }
        R.ap[q] = o;
    }
    return acc;
}
// PCGStep2 (:446-489) for the thread's pixels; pixels within two of the tile's border publish their z for the neighbouring tiles' rings
template <class T, int KMAX>
SFS_HD void sfsLoopStep2(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT, T alpha, int usePre, unsigned long long* zLL, unsigned tag, T& accRz, T& accQ) {
    const int PW = V.PW();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            const T d = R.delta[q] + alpha * V.sp[(ty + 2) * PW + tx + 2];
            const T rr = R.r[q] - alpha * R.ap[q];
            const T zz = usePre ? R.pre[q] * rr : rr;
            R.delta[q] = d; R.r[q] = rr; R.z[q] = zz;
            accRz += zz * rr;
            accQ += T(0.5) * (d * (rr + R.b[q]));
            if (ty < 2 || ty >= V.th - 2 || tx < 2 || tx >= V.tw - 2) sfsEdgeStore<T>(zLL, gi, zz, tag);
        }
    }
}
// ---- residual-reset iteration (LM, every residual_reset_period-th iteration; :1077-1083) in three pieces:
// PCGStep2_1stHalf (:491-503): delta += alpha p; the new delta goes into the tile's delta plane and, for pixels within two of the border,
// into the edge words for the neighbouring tiles' rings
template <class T, int KMAX>
SFS_HD void sfsLoopResetFirstHalf(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT, T alpha, unsigned long long* dLL, unsigned tag) {
    const int PW = V.PW();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            const int cc = (ty + 2) * PW + tx + 2;
            const T d = R.delta[q] + alpha * V.sp[cc];
            R.delta[q] = d; V.sd[cc] = d;
            if (ty < 2 || ty >= V.th - 2 || tx < 2 || tx >= V.tw - 2) sfsEdgeStore<T>(dLL, gi, d, tag);
        }
    }
}
// the ring of the delta plane from what the neighbouring tiles published (each word is awaited by its tag: no barrier in between)
template <class T>
SFS_HD void sfsLoopFetchRing(const SfsData<T>& P, const SfsTileView<T>& V, int tid, int NT, T* plane, const unsigned long long* vLL, unsigned tag) {
    const int PW = V.PW(), PH = V.PH();
    const int nRing = 4 * PW + 4 * V.th;
    for (int j = tid; j < nRing; j += NT) {
        int ly, lx;
        if (j < 2 * PW) { ly = j >= PW ? 1 : 0; lx = j - ly * PW; }
        else if (j < 4 * PW) { const int jj = j - 2 * PW; ly = jj >= PW ? 1 : 0; lx = jj - ly * PW; ly += PH - 2; }
        else { const int jj = j - 4 * PW; ly = 2 + (jj >> 2); const int s = jj & 3; lx = s < 2 ? s : PW - 4 + s; }
        const int gx = V.x0 - 2 + lx, gy = V.y0 - 2 + ly;
        if (sfsInImage(P, gx, gy)) plane[ly * PW + lx] = sfsEdgeLoad<T>(vLL, (long)gy * P.W + gx, tag);
    }
}
// A delta out of the row values of the delta plane, then PCGStep2_2ndHalf (:505-534): r = b - A delta, z = pre r, r.z and q
template <class T, int KMAX>
SFS_HD void sfsLoopResetSecondHalf(const SfsData<T>& P, const SfsTileView<T>& V, SfsLoopRegs<T, KMAX>& R, int tid, int NT, int usePre, unsigned long long* zLL, unsigned tag, T& accRz, T& accQ) {
    const int PW = V.PW(), RW = V.RW();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            T ad = T(0);
            if (V.sfl[(ty + 1) * RW + tx + 1] & kSfsIncluded) {
                const T di = V.sd[(ty + 2) * PW + tx + 2];
                ad = sfsTileGather(P, V, ty, tx, P.wp * (P.wp * di));
                ad += R.ctc[q] * di;
            }
            const T rr = R.b[q] - ad;
            const T zz = usePre ? R.pre[q] * rr : rr;
            R.r[q] = rr; R.z[q] = zz;
            accRz += zz * rr;
            accQ += T(0.5) * (R.delta[q] * (rr + R.b[q]));
            if (ty < 2 || ty >= V.th - 2 || tx < 2 || tx >= V.tw - 2) sfsEdgeStore<T>(zLL, gi, zz, tag);
        }
    }
}
// ---- computeModelCost (:666-678) at the end of the loop: delta into the delta plane (+ tile-edge publish, as in the reset iteration), then
// 1/2 sum (F + J delta)^2 over the rows centred at the thread's included pixels, J delta taken from the row values of the delta plane
template <class T, int KMAX>
SFS_HD void sfsLoopPublishDelta(const SfsData<T>& P, const SfsTileView<T>& V, const SfsLoopRegs<T, KMAX>& R, int tid, int NT, unsigned long long* dLL, unsigned tag) {
    const int PW = V.PW();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            V.sd[(ty + 2) * PW + tx + 2] = R.delta[q];
            if (ty < 2 || ty >= V.th - 2 || tx < 2 || tx >= V.tw - 2) sfsEdgeStore<T>(dLL, gi, R.delta[q], tag);
        }
    }
}
template <class T, int KMAX>
SFS_HD T sfsLoopModelCost(const SfsData<T>& P, const SfsTileView<T>& V, const SfsLoopRegs<T, KMAX>& R, int tid, int NT) {
    const int RW = V.RW(), cells = V.cells();
    T acc = T(0);
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (!sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) continue;
        const int cc = (ty + 1) * RW + tx + 1;
        const unsigned f = V.sfl[cc];
        if (!(f & kSfsIncluded)) continue;
        T e = T(0);
        { const T r = sfsFP(P, gi) + P.wp * R.delta[q]; e += r * r; }
        if (f & kSfsInner) {
            const T Bc = P.B[gi];
            const T r0 = P.wg * ((Bc - P.B[gi + 1]) * T(P.eR[gi])) + V.sy[cc];
            const T r1 = P.wg * ((Bc - P.B[gi + P.W]) * T(P.eC[gi])) + V.sy[cells + cc];
            e += r0 * r0; e += r1 * r1;
        }
        if (f & kSfsFlat) {
            const int lx = tx + 1, ly = ty + 1;   // the pixel's row cell; factor tables are indexed like the vector planes
            const T ax[3] = { V.sAx[lx], V.sAx[lx + 1], V.sAx[lx + 2] }, ay[3] = { V.sAy[ly], V.sAy[ly + 1], V.sAy[ly + 2] };
            T fs[3] = { T(0), T(0), T(0) };
            SFS_UNROLL for (int m = 0; m < 5; ++m) {
                const T xm = P.X[gi + (long)sfsDy(2, m) * P.W + sfsDx(2, m)];
                fs[0] += sfsCoefFromA(P, ax[1 + sfsDx(2, m)], m) * xm;
                fs[1] += sfsCoefFromA(P, ay[1 + sfsDy(2, m)], m) * xm;
                fs[2] += sfsCoefFromA(P, T(1), m) * xm;
            }
            SFS_UNROLL for (int comp = 0; comp < 3; ++comp) { const T r = fs[comp] + V.sy[(2 + comp) * cells + cc]; e += r * r; }
        }
        acc += T(0.5) * e;
    }
    return acc;
}
// epilogue: the last PCGStep3 and the state back to the plan's vectors
template <class T, int KMAX>
SFS_HD void sfsLoopStore(const SfsData<T>& P, const SfsTileView<T>& V, const SfsLoopRegs<T, KMAX>& R, int tid, int NT, bool stepped, T beta,
                         T* delta, T* r, T* z, T* p, T* Ap, T* X = nullptr, T* prevX = nullptr,   // X != nullptr: savePreviousUnknowns (prevX) + PCGLinearUpdate too
                         T* preOut = nullptr, T* ctcOut = nullptr, T* bOut = nullptr) {            // head inside the kernel: its vectors go back as well
    const int PW = V.PW();
    SFS_UNROLL for (int q = 0; q < KMAX; ++q) {
        int ty, tx; long gi;
        if (sfsLoopPixel(P, V, tid + q * NT, ty, tx, gi)) {
            T pp = V.sp[(ty + 2) * PW + tx + 2];
            if (stepped) pp = R.z[q] + beta * pp;
            p[gi] = pp; delta[gi] = R.delta[q]; r[gi] = R.r[q]; z[gi] = R.z[q]; Ap[gi] = R.ap[q];
            if (X) { const T x = X[gi]; if (prevX) prevX[gi] = x; X[gi] = x + R.delta[q]; }
            if (preOut) { preOut[gi] = R.pre[q]; ctcOut[gi] = R.ctc[q]; bOut[gi] = R.b[q]; }
        }
    }
}

}  // namespace optb
