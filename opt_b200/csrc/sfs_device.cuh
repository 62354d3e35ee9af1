// sfs_device.cuh -- per-pixel math of the hand-written shape_from_shading kernel set (energy_sfs.cu).
//
// Energy (reference examples/shape_from_shading/shape_from_shading.t:1-95; specs/shape_from_shading.t here), per pixel c
// of a W x H depth map X with measured depth D, intensity Im and two uint8 edge masks:
//     r_p(c)    = sqrt(w_p) (X(c) - D(c))                                   if D(c) > 0
//     r_gR(c)   = sqrt(w_g) edgeMaskR(c) (B(c) - B(c+(1,0)))                if c is at least one pixel inside the image
//     r_gC(c)   = sqrt(w_g) edgeMaskC(c) (B(c) - B(c+(0,1)))                if c is at least one pixel inside the image
//     r_s,k(c)  = sqrt(w_s) (4 P_k(c) - P_k(c-(1,0)) - P_k(c-(0,1)) - P_k(c+(1,0)) - P_k(c+(0,1))),  k = x,y,z   if flat(c)
//   B(c) = shading(c) - intensity(c): second-order spherical-harmonics shading of the normal built from X(c), X(c-(1,0)),
//   X(c-(0,1)) (zero unless c is inside and those three depths are valid); P(c) = back-projected point; flat(c): the five
//   depths are valid and differ by less than 0.01.  Exclude(D(c) <= 0).  No UsePreconditioner => false (o.t:214).
// What the reference generates for this spec: B_I and flat are ComputedArrays refreshed by kernels.precompute
// (solverGPUGaussNewton.t:607-614, o.t:1007-1040); J^T J p / J^T F / cost gather over the residuals that read X(c)
// (o.t:1996-2172) with the chain rule through B_I (o.t:913-925).
//
// Here, once per linearisation: B and its three partials (dual numbers with three slots), then the FIVE coefficients of
// every r_gR / r_gC row (10 floats per pixel); r_p and r_s rows are affine in X with coefficients that depend on the pixel
// coordinates only and are rebuilt on the fly.  Each PCG step is then ONE gather kernel: for unknown u, every row that reads
// X(u) is visited (5 r_gR centres, 5 r_gC centres, 5 r_s centres, r_p), y = row . p is formed and credited with the row's
// coefficient for u.  Rows centred at excluded pixels do not count in the cost but do feed these gathers (SURVEY.md 8a
// note 2); a row never has a non-zero coefficient on an excluded unknown (every B / flat term requires valid depths).
//
// Everything in this header is plain per-pixel code callable from host and device: tests/host/sfs_host.cpp runs the same
// functions on the CPU against the oracle (cost, J^T F, diag J^T J, J^T J x), so the GPU kernels in energy_sfs.cu only add
// the one-thread-per-pixel loops and the reductions.
#pragma once
#if defined(__CUDACC__)
#define SFS_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SFS_HD inline
#endif

namespace optb {

template <class T>
struct SfsData {
    int W, H;
    T wp, ws, wg, fx, fy, ux, uy, L[10];          // sqrt of the three weights (spec lines 8-10); L[1..9] lighting
    const T* X; const T* D; const T* Im; const unsigned char* eR; const unsigned char* eC;
    // plan-owned, refreshed once per linearisation (precompute):
    T* B;                // B(c)
    T* dB;               // [3][N]: dB/dX(c), dB/dX(c-(1,0)), dB/dX(c-(0,1))
    T* cR; T* cC;        // [5][N] each: coefficients of the r_gR / r_gC row centred at the pixel, dependency order below
    unsigned char* fl;   // bit 0: D > 0 (included), bit 1: at least one pixel inside the image, bit 2: flat
};
// dependency order (offsets from the row's centre):
//   r_gR: (0,0) (-1,0) (0,-1) (1,0) (1,-1)      r_gC: (0,0) (-1,0) (0,-1) (0,1) (-1,1)      r_s: (0,0) (-1,0) (0,-1) (1,0) (0,1)
// as functions of (row type t: 0 = r_gR, 1 = r_gC, 2 = r_s; dependency k): with the loops below fully unrolled these fold to
// constants (a by-value table indexed at run time would live in local memory)
SFS_HD int sfsDx(int t, int k) { return t == 0 ? (k == 1 ? -1 : (k >= 3 ? 1 : 0)) : (t == 1 ? ((k == 1 || k == 4) ? -1 : 0) : (k == 1 ? -1 : (k == 3 ? 1 : 0))); }
SFS_HD int sfsDy(int t, int k) { return t == 0 ? ((k == 2 || k == 4) ? -1 : 0) : (t == 1 ? (k == 2 ? -1 : (k >= 3 ? 1 : 0)) : (k == 2 ? -1 : (k == 4 ? 1 : 0))); }
#if defined(__CUDACC__)
#define SFS_UNROLL _Pragma("unroll")
#else
#define SFS_UNROLL
#endif

constexpr unsigned kSfsIncluded = 1u, kSfsInner = 2u, kSfsFlat = 4u;

// ---- three-slot dual number for dB
template <class T> struct SfsDual { T v, d0, d1, d2; };
template <class T> SFS_HD SfsDual<T> sfsC(T c) { return SfsDual<T>{ c, T(0), T(0), T(0) }; }
template <class T> SFS_HD SfsDual<T> operator+(SfsDual<T> a, SfsDual<T> b) { return SfsDual<T>{ a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2 }; }
template <class T> SFS_HD SfsDual<T> operator-(SfsDual<T> a, SfsDual<T> b) { return SfsDual<T>{ a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2 }; }
template <class T> SFS_HD SfsDual<T> operator-(SfsDual<T> a) { return SfsDual<T>{ -a.v, -a.d0, -a.d1, -a.d2 }; }
template <class T> SFS_HD SfsDual<T> operator*(SfsDual<T> a, SfsDual<T> b) { return SfsDual<T>{ a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2 }; }
template <class T> SFS_HD SfsDual<T> operator*(T a, SfsDual<T> b) { return SfsDual<T>{ a * b.v, a * b.d0, a * b.d1, a * b.d2 }; }
template <class T> SFS_HD SfsDual<T> operator*(SfsDual<T> a, T b) { return b * a; }
template <class T> SFS_HD SfsDual<T> operator/(SfsDual<T> a, T b) { return SfsDual<T>{ a.v / b, a.d0 / b, a.d1 / b, a.d2 / b }; }
template <class T> SFS_HD SfsDual<T> operator+(T a, SfsDual<T> b) { return SfsDual<T>{ a + b.v, b.d0, b.d1, b.d2 }; }
SFS_HD float sfsSqrt(float x) { return sqrtf(x); }
SFS_HD double sfsSqrt(double x) { return sqrt(x); }
SFS_HD float sfsAbs(float x) { return fabsf(x); }
SFS_HD double sfsAbs(double x) { return fabs(x); }
template <class T> SFS_HD SfsDual<T> sfsRsqrt(SfsDual<T> a) {   // 1 / sqrt(a)
    const T s = sfsSqrt(a.v), v = T(1) / s, k = T(-0.5) * v / a.v;
    return SfsDual<T>{ v, k * a.d0, k * a.d1, k * a.d2 };
}

template <class T> SFS_HD bool sfsInside1(const SfsData<T>& P, int x, int y) { return x - 1 >= 0 && x + 1 < P.W && y - 1 >= 0 && y + 1 < P.H; }   // InBoundsExpanded(0,0,1)
template <class T> SFS_HD bool sfsInImage(const SfsData<T>& P, int x, int y) { return x >= 0 && x < P.W && y >= 0 && y < P.H; }
template <class T> SFS_HD bool sfsHasDepth(const SfsData<T>& P, int x, int y) { return P.D[(long)y * P.W + x] > T(0); }
template <class T> SFS_HD long sfsN(const SfsData<T>& P) { return (long)P.W * P.H; }

// B(c) = shading - intensity and its partials w.r.t. X(c), X(c-(1,0)), X(c-(0,1))  (spec lines 34-66)
template <class T> SFS_HD SfsDual<T> sfsB(const SfsData<T>& P, int x, int y) {
    if (!sfsInImage(P, x, y) || !sfsInside1(P, x, y)) return sfsC(T(0));
    if (!(sfsHasDepth(P, x - 1, y) && sfsHasDepth(P, x, y) && sfsHasDepth(P, x, y - 1))) return sfsC(T(0));
    const long c = (long)y * P.W + x;
    const SfsDual<T> x00{ P.X[c], T(1), T(0), T(0) }, xm0{ P.X[c - 1], T(0), T(1), T(0) }, x0m{ P.X[c - P.W], T(0), T(0), T(1) };
    const T i = T(x), j = T(y);
    SfsDual<T> nx = x0m * (x00 - xm0) / P.fy;
    SfsDual<T> ny = xm0 * (x00 - x0m) / P.fx;
    SfsDual<T> nz = (nx * ((P.ux - i) / P.fx)) + (ny * ((P.uy - j) / P.fy)) - (xm0 * x0m / (P.fx * P.fy));
    const SfsDual<T> len2 = nx * nx + ny * ny + nz * nz;
    if (len2.v > T(0)) { const SfsDual<T> inv = sfsRsqrt(len2); nx = inv * nx; ny = inv * ny; nz = inv * nz; }
    const SfsDual<T> sh = P.L[1] + (P.L[2] * ny + P.L[3] * nz + P.L[4] * nx + P.L[5] * (nx * ny) + P.L[6] * (ny * nz)
                                    + P.L[7] * (-(nx * nx) - ny * ny + T(2) * (nz * nz)) + P.L[8] * (nz * nx) + P.L[9] * (nx * nx - ny * ny));
    const T I = P.Im[c] * T(0.5) + T(0.25) * (P.Im[c - 1] + P.Im[c - P.W]);
    return SfsDual<T>{ sh.v - I, sh.d0, sh.d1, sh.d2 };
}

// flag byte of the in-image pixel (x, y)
template <class T> SFS_HD unsigned sfsFlagsAt(const SfsData<T>& P, int x, int y) {
    const long c = (long)y * P.W + x;
    unsigned f = 0;
    if (sfsHasDepth(P, x, y)) f |= kSfsIncluded;
    if (sfsInside1(P, x, y)) {
        f |= kSfsInner;
        const T v = P.X[c], thre = T(0.01);
        const bool flat = sfsHasDepth(P, x, y) && sfsHasDepth(P, x, y - 1) && sfsHasDepth(P, x, y + 1) && sfsHasDepth(P, x - 1, y) && sfsHasDepth(P, x + 1, y)
            && sfsAbs(v - P.X[c - P.W]) < thre && sfsAbs(v - P.X[c + P.W]) < thre && sfsAbs(v - P.X[c - 1]) < thre && sfsAbs(v - P.X[c + 1]) < thre;
        if (flat) f |= kSfsFlat;
    }
    return f;
}
// precompute, pass 1: B, dB and the flag byte of pixel (x, y)
template <class T> SFS_HD void sfsPrecomputeB(const SfsData<T>& P, int x, int y) {
    const long c = (long)y * P.W + x, N = sfsN(P);
    const SfsDual<T> b = sfsB(P, x, y);
    P.B[c] = b.v; P.dB[c] = b.d0; P.dB[N + c] = b.d1; P.dB[2 * N + c] = b.d2;
    P.fl[c] = (unsigned char)sfsFlagsAt(P, x, y);
}
// precompute, pass 2 (needs pass 1 of the right and lower neighbours): coefficients of the r_gR / r_gC rows centred at (x, y)
template <class T> SFS_HD void sfsBuildRows(const SfsData<T>& P, int x, int y) {
    const long c = (long)y * P.W + x, N = sfsN(P);
    T r[5] = { T(0), T(0), T(0), T(0), T(0) }, q[5] = { T(0), T(0), T(0), T(0), T(0) };
    if (P.fl[c] & kSfsInner) {
        const T d0 = P.dB[c], d1 = P.dB[N + c], d2 = P.dB[2 * N + c];
        {   // B(c) - B(c+(1,0)):  B(c+(1,0)) reads X(c+(1,0)) [slot 0], X(c) [slot 1], X(c+(1,-1)) [slot 2]
            const T s = P.wg * T(P.eR[c]); const long n = c + 1;
            r[0] = s * (d0 - P.dB[N + n]); r[1] = s * d1; r[2] = s * d2; r[3] = -s * P.dB[n]; r[4] = -s * P.dB[2 * N + n];
        }
        {   // B(c) - B(c+(0,1)):  B(c+(0,1)) reads X(c+(0,1)) [slot 0], X(c+(-1,1)) [slot 1], X(c) [slot 2]
            const T s = P.wg * T(P.eC[c]); const long n = c + P.W;
            q[0] = s * (d0 - P.dB[2 * N + n]); q[1] = s * d1; q[2] = s * d2; q[3] = -s * P.dB[n]; q[4] = -s * P.dB[N + n];
        }
    }
    SFS_UNROLL for (int k = 0; k < 5; ++k) { P.cR[(long)k * N + c] = r[k]; P.cC[(long)k * N + c] = q[k]; }
}

// ---- residual values of the rows centred at (x, y) (for J^T F, cost, model cost)
template <class T> SFS_HD T sfsFP(const SfsData<T>& P, long c) { return (P.fl[c] & kSfsIncluded) ? P.wp * (P.X[c] - P.D[c]) : T(0); }
template <class T> SFS_HD T sfsFR(const SfsData<T>& P, long c) { return (P.fl[c] & kSfsInner) ? P.wg * ((P.B[c] - P.B[c + 1]) * T(P.eR[c])) : T(0); }
template <class T> SFS_HD T sfsFC(const SfsData<T>& P, long c) { return (P.fl[c] & kSfsInner) ? P.wg * ((P.B[c] - P.B[c + P.W]) * T(P.eC[c])) : T(0); }
// coefficient of the k-th dependency (order sfsDepsS) in component comp of the r_s row centred at (x, y); zero row unless flat
// back-projection factor of pixel column qx / row qy (P(c) = (ax d, ay d, d)), and the row coefficient built from it
template <class T> SFS_HD T sfsAx(const SfsData<T>& P, int qx) { return (T(qx) - P.ux) / P.fx; }
template <class T> SFS_HD T sfsAy(const SfsData<T>& P, int qy) { return (T(qy) - P.uy) / P.fy; }
template <class T> SFS_HD T sfsCoefFromA(const SfsData<T>& P, T a, int k) { return k == 0 ? P.ws * (T(4) * a) : -(P.ws * a); }
template <class T> SFS_HD T sfsCoefS(const SfsData<T>& P, int x, int y, int comp, int k) {
    const int qx = x + sfsDx(2, k), qy = y + sfsDy(2, k);
    const T a = comp == 0 ? sfsAx(P, qx) : (comp == 1 ? sfsAy(P, qy) : T(1));
    return sfsCoefFromA(P, a, k);
}
template <class T> SFS_HD T sfsFSWithFlag(const SfsData<T>& P, int x, int y, int comp);
template <class T> SFS_HD T sfsFS(const SfsData<T>& P, int x, int y, int comp) {   // value of r_s,comp centred at (x, y)
    const long c = (long)y * P.W + x;
    if (!(P.fl[c] & kSfsFlat)) return T(0);
    return sfsFSWithFlag(P, x, y, comp);
}
template <class T> SFS_HD T sfsFSWithFlag(const SfsData<T>& P, int x, int y, int comp) {   // the same for a centre the caller knows to be flat
    const long c = (long)y * P.W + x;
    // the five coefficients need the factor of three columns (comp 0) or three rows (comp 1): 3 divisions instead of one per coefficient
    const T a3[3] = { comp == 0 ? sfsAx(P, x - 1) : (comp == 1 ? sfsAy(P, y - 1) : T(1)), comp == 0 ? sfsAx(P, x) : (comp == 1 ? sfsAy(P, y) : T(1)),
                      comp == 0 ? sfsAx(P, x + 1) : (comp == 1 ? sfsAy(P, y + 1) : T(1)) };
    T s = T(0);
    SFS_UNROLL for (int k = 0; k < 5; ++k) s += sfsCoefFromA(P, a3[1 + (comp == 1 ? sfsDy(2, k) : sfsDx(2, k))], k) * P.X[c + (long)sfsDy(2, k) * P.W + sfsDx(2, k)];
    return s;
}

// ---- gathers for unknown u = (x, y).  `vec` is the vector the rows are applied to (p, delta) or nullptr for J^T F,
// where the row's residual value takes the place of row . vec.  MODE 0: sum_rows coef(row,u) * (row . vec)   [J^T J vec]
//                                                                  MODE 1: sum_rows coef(row,u) * F(row)       [J^T F]
//                                                                  MODE 2: sum_rows coef(row,u)^2              [diag J^T J]
template <class T, int MODE> SFS_HD T sfsGather(const SfsData<T>& P, int x, int y, const T* vec) {
    const long N = sfsN(P), u = (long)y * P.W + x;
    if (!(P.fl[u] & kSfsIncluded)) return T(0);
    T acc = T(0);
    // r_p: one dependency, the centre itself
    {
        const T a = P.wp;
        acc += MODE == 0 ? a * (a * vec[u]) : (MODE == 1 ? a * sfsFP(P, u) : a * a);
    }
    // r_gR and r_gC: u is dependency k of the row centred at c = u - offset_k
    SFS_UNROLL for (int t = 0; t < 2; ++t) {
        const T* coef = t == 0 ? P.cR : P.cC;
        SFS_UNROLL for (int k = 0; k < 5; ++k) {
            const int cx = x - sfsDx(t, k), cy = y - sfsDy(t, k);
            if (!sfsInImage(P, cx, cy)) continue;
            const long c = (long)cy * P.W + cx;
            if (!(P.fl[c] & kSfsInner)) continue;
            const T a = coef[(long)k * N + c];
            if (a == T(0)) continue;
            if (MODE == 2) { acc += a * a; continue; }
            T yv;
            if (MODE == 1) yv = t == 0 ? sfsFR(P, c) : sfsFC(P, c);
            else {
                yv = T(0);
                SFS_UNROLL for (int m = 0; m < 5; ++m) {
                    const T am = coef[(long)m * N + c];
                    if (am != T(0)) yv += am * vec[c + (long)sfsDy(t, m) * P.W + sfsDx(t, m)];   // a non-zero coefficient implies an in-image, included dependency
                }
            }
            acc += a * yv;
        }
    }
    // r_s (three components): u is dependency k of the row centred at c = u - offset_k, if that centre is flat.  Dependency k of that row
    // IS u, so its coefficient is built from u's own back-projection factors: two divisions per unknown, not one per (row, component)
    {
        const T aU[3] = { sfsAx(P, x), sfsAy(P, y), T(1) };
        SFS_UNROLL for (int k = 0; k < 5; ++k) {
            const int cx = x - sfsDx(2, k), cy = y - sfsDy(2, k);
            if (!sfsInImage(P, cx, cy)) continue;
            const long c = (long)cy * P.W + cx;
            if (!(P.fl[c] & kSfsFlat)) continue;
            SFS_UNROLL for (int comp = 0; comp < 3; ++comp) {
                const T a = sfsCoefFromA(P, aU[comp], k);
                if (MODE == 2) { acc += a * a; continue; }
                T yv;
                if (MODE == 1) yv = sfsFS(P, cx, cy, comp);
                else {
                    yv = T(0);
                    SFS_UNROLL for (int m = 0; m < 5; ++m) yv += sfsCoefS(P, cx, cy, comp, m) * vec[c + (long)sfsDy(2, m) * P.W + sfsDx(2, m)];
                }
                acc += a * yv;
            }
        }
    }
    return acc;
}

// 1/2 sum of squares of the six rows centred at (x, y); MODEL: rows are F + J delta (computeModelCost).  Rows centred at
// excluded pixels do not count (computeCost skips them, o.t:2375-2385).
template <class T, bool MODEL> SFS_HD T sfsCostAt(const SfsData<T>& P, int x, int y, const T* delta) {
    const long N = sfsN(P), c = (long)y * P.W + x;
    if (!(P.fl[c] & kSfsIncluded)) return T(0);
    T e = T(0);
    { T r = sfsFP(P, c); if (MODEL) r += P.wp * delta[c]; e += r * r; }
    if (P.fl[c] & kSfsInner) {
        SFS_UNROLL for (int t = 0; t < 2; ++t) {
            const T* coef = t == 0 ? P.cR : P.cC;
            T r = t == 0 ? sfsFR(P, c) : sfsFC(P, c);
            if (MODEL) { SFS_UNROLL for (int m = 0; m < 5; ++m) { const T am = coef[(long)m * N + c]; if (am != T(0)) r += am * delta[c + (long)sfsDy(t, m) * P.W + sfsDx(t, m)]; } }
            e += r * r;
        }
    }
    if (P.fl[c] & kSfsFlat) {
        SFS_UNROLL for (int comp = 0; comp < 3; ++comp) {
            T r = sfsFS(P, x, y, comp);
            if (MODEL) { SFS_UNROLL for (int m = 0; m < 5; ++m) r += sfsCoefS(P, x, y, comp, m) * delta[c + (long)sfsDy(2, m) * P.W + sfsDx(2, m)]; }
            e += r * r;
        }
    }
    return T(0.5) * e;
}

}  // namespace optb
