// energies_generic.cu -- energy specifications for the generic cached-Jacobian engine (generic_engine.cuh).
//
// Each struct states one reference spec's residuals the way the .t file does (file:line cited per struct);
// the engine differentiates them on the device.  These are what a general front end would generate.
// Graph index arrays are int32 per vertex slot, the edge count comes from a host int* (util.t:677-682);
// float Params stay fp32 and are cast on use (o.t:1699).
#include "generic_engine.cuh"

namespace optb {

template <class S> struct G3 { S x, y, z; };
template <class S> struct G2 { S x, y; };
// lib.t:92-95 Rotate2D
template <class S> __device__ __forceinline__ G2<S> geRotate2D(const S& a, const G2<S>& v) {
    const S c = geCos(a), s = geSin(a);
    return { c * v.x - s * v.y, s * v.x + c * v.y };
}
// lib.t:77-91 Rotate3D
template <class S> __device__ __forceinline__ G3<S> geRotate3D(const G3<S>& a, const G3<S>& v) {
    const S ca = geCos(a.x), cb = geCos(a.y), cg = geCos(a.z), sa = geSin(a.x), sb = geSin(a.y), sg = geSin(a.z);
    const S m0 = cg * cb, m1 = -sg * ca + cg * sb * sa, m2 = sg * sa + cg * sb * ca;
    const S m3 = sg * cb, m4 = cg * ca + sg * sb * sa, m5 = -cg * sa + sg * sb * ca;
    const S m6 = -sb, m7 = cb * sa, m8 = cb * ca;
    return { m0 * v.x + m1 * v.y + m2 * v.z, m3 * v.x + m4 * v.y + m5 * v.z, m6 * v.x + m7 * v.y + m8 * v.z };
}

// ============================================================ image_warping.t:1-23 (cross-check of the hand-written kernels)
template <class T>
struct ImageWarpingSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; int W, H; const T *UrShape, *Constraints, *Mask; T wfit, wreg; };
    int W, H; bool usePreconditioner = true, usesGraph = false;
    ImageWarpingSpec(const unsigned* d) : W((int)d[0]), H((int)d[1]) {}
    static const char* name() { return "image_warping (generic engine)"; }
    int numArrays() const { return 2; }
    void describe(int a, long& el, int& ch) const { el = (long)W * H; ch = a == 0 ? 2 : 1; }
    void bind(void** q, Params& P) const {
        P.W = W; P.H = H; P.U[0].data = (T*)q[0]; P.U[1].data = (T*)q[1];
        P.UrShape = (const T*)q[2]; P.Constraints = (const T*)q[3]; P.Mask = (const T*)q[4];
        P.wfit = (T) * (const float*)q[5]; P.wreg = (T) * (const float*)q[6];
    }
    int numGroups(const Params&) const { return 5; }
    GeGroup group(const Params&, int g) const { return GeGroup{ (long)W * H, 2, g < 4 ? 5 : 2, 1, 0 }; }
    __device__ static bool excluded(const Params& P, int, long e) { return !(P.Mask[e] == T(0)); }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        const int x = (int)(i % P.W), y = (int)(i / P.W);
        if (g < 4) {
            const int nx = x + (g == 0 ? 1 : (g == 1 ? -1 : 0)), ny = y + (g == 2 ? 1 : (g == 3 ? -1 : 0));
            const bool inb = nx >= 0 && nx < P.W && ny >= 0 && ny < P.H;
            const long j = inb ? (long)ny * P.W + nx : i;
            const bool valid = inb && P.Mask[j] == T(0) && P.Mask[i] == T(0);
            const S o0x = X(0, i, 0), o0y = X(0, i, 1), o1x = X(0, j, 0), o1y = X(0, j, 1), a = X(1, i, 0);
            const G2<S> d = { S(P.UrShape[2 * i] - P.UrShape[2 * j]), S(P.UrShape[2 * i + 1] - P.UrShape[2 * j + 1]) };
            const G2<S> r = geRotate2D(a, d);
            const S ex = P.wreg * ((o0x - o1x) - r.x), ey = P.wreg * ((o0y - o1y) - r.y);
            out[0] = valid ? ex : S(T(0)); out[1] = valid ? ey : S(T(0));
        } else {
            const bool valid = P.Constraints[2 * i] >= T(0) && P.Constraints[2 * i + 1] >= T(0);
            const S ex = X(0, i, 0) - P.Constraints[2 * i], ey = X(0, i, 1) - P.Constraints[2 * i + 1];
            out[0] = P.wfit * (valid ? ex : S(T(0))); out[1] = P.wfit * (valid ? ey : S(T(0)));
        }
    }
};

// ============================================================ arap_mesh_deformation.t:1-18
template <class T>
struct ArapSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; long N; int nEdges; const int *v0, *v1; const T *UrShape, *Constraints; T wfit, wreg; };
    long N; bool usePreconditioner = true, usesGraph = true;
    ArapSpec(const unsigned* d) : N(d[0]) {}
    static const char* name() { return "arap_mesh_deformation"; }
    int numArrays() const { return 2; }
    void describe(int, long& el, int& ch) const { el = N; ch = 3; }
    void bind(void** q, Params& P) const {
        P.N = N; P.wfit = (T) * (const float*)q[0]; P.wreg = (T) * (const float*)q[1];
        P.U[0].data = (T*)q[2]; P.U[1].data = (T*)q[3]; P.UrShape = (const T*)q[4]; P.Constraints = (const T*)q[5];
        P.nEdges = *(const int*)q[6]; P.v0 = (const int*)q[7]; P.v1 = (const int*)q[8];
    }
    int numGroups(const Params&) const { return 2; }
    GeGroup group(const Params& P, int g) const { return g == 0 ? GeGroup{ N, 3, 3, 1, 0 } : GeGroup{ (long)P.nEdges, 3, 9, 0, 0 }; }
    __device__ static bool excluded(const Params&, int, long) { return false; }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        if (g == 0) {
            const bool valid = P.Constraints[3 * i] >= T(-999999.9);   // channel 0 only (spec line 12)
            for (int c = 0; c < 3; ++c) { const S e = P.wfit * (X(0, i, c) - P.Constraints[3 * i + c]); out[c] = valid ? e : S(T(0)); }
        } else {
            const long a = P.v0[i], b = P.v1[i];
            const G3<S> ang = { X(1, a, 0), X(1, a, 1), X(1, a, 2) };
            const G3<S> d = { S(P.UrShape[3 * a] - P.UrShape[3 * b]), S(P.UrShape[3 * a + 1] - P.UrShape[3 * b + 1]), S(P.UrShape[3 * a + 2] - P.UrShape[3 * b + 2]) };
            const G3<S> r = geRotate3D(ang, d);
            out[0] = P.wreg * ((X(0, a, 0) - X(0, b, 0)) - r.x);
            out[1] = P.wreg * ((X(0, a, 1) - X(0, b, 1)) - r.y);
            out[2] = P.wreg * ((X(0, a, 2) - X(0, b, 2)) - r.z);
        }
    }
};

// ============================================================ poisson_image_editing.t:1-13  (UsePreconditioner(false))
template <class T>
struct PoissonSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; int W, H; const T *Tg, *M; };
    int W, H; bool usePreconditioner = false, usesGraph = false;
    PoissonSpec(const unsigned* d) : W((int)d[0]), H((int)d[1]) {}
    static const char* name() { return "poisson_image_editing"; }
    int numArrays() const { return 1; }
    void describe(int, long& el, int& ch) const { el = (long)W * H; ch = 4; }
    void bind(void** q, Params& P) const { P.W = W; P.H = H; P.U[0].data = (T*)q[0]; P.Tg = (const T*)q[1]; P.M = (const T*)q[2]; }
    int numGroups(const Params&) const { return 4; }
    GeGroup group(const Params&, int) const { return GeGroup{ (long)W * H, 4, 8, 1, 0 }; }
    __device__ static bool excluded(const Params& P, int, long e) { return !(P.M[e] == T(0)); }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        const int x = (int)(i % P.W), y = (int)(i / P.W);
        const int nx = x + (g == 0 ? 1 : (g == 1 ? -1 : 0)), ny = y + (g == 2 ? 1 : (g == 3 ? -1 : 0));
        const bool inb = nx >= 0 && nx < P.W && ny >= 0 && ny < P.H;
        const long j = inb ? (long)ny * P.W + nx : i;
        for (int c = 0; c < 4; ++c) {
            const S e = (X(0, i, c) - X(0, j, c)) - (P.Tg[4 * i + c] - P.Tg[4 * j + c]);
            out[c] = inb ? e : S(T(0));
        }
    }
};

// ============================================================ tests/minimal/laplacian.t:1-8  (no UsePreconditioner call: default false, o.t:214;
// default bounds rule o.t:1930-1934 zeroes a residual whose stencil leaves the image)
template <class T>
struct LaplacianSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; int W, H; const T* A; };
    int W, H; bool usePreconditioner = false, usesGraph = false;
    LaplacianSpec(const unsigned* d) : W((int)d[0]), H((int)d[1]) {}
    static const char* name() { return "laplacian"; }
    int numArrays() const { return 1; }
    void describe(int, long& el, int& ch) const { el = (long)W * H; ch = 1; }
    void bind(void** q, Params& P) const { P.W = W; P.H = H; P.U[0].data = (T*)q[0]; P.A = (const T*)q[1]; }
    int numGroups(const Params&) const { return 3; }
    GeGroup group(const Params&, int) const { return GeGroup{ (long)W * H, 1, 2, 1, 0 }; }
    __device__ static bool excluded(const Params&, int, long) { return false; }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        const int x = (int)(i % P.W), y = (int)(i / P.W);
        if (g == 0) out[0] = T(.2) * (X(0, i, 0) - P.A[i]);
        else if (g == 1) { const bool inb = x + 1 < P.W; const long j = inb ? i + 1 : i; const S e = X(0, i, 0) - X(0, j, 0); out[0] = inb ? e : S(T(0)); }
        else { const bool inb = y + 1 < P.H; const long j = inb ? i + P.W : i; const S e = X(0, i, 0) - X(0, j, 0); out[0] = inb ? e : S(T(0)); }
    }
};

// ============================================================ tests/minimal_graph_only/curveFitting.t:1-9 ; dims = {N, U}
// The unknown index space U has no centred residual: the reference adds a zero stand-in energy
// (o.t:1938-1982) -- group 0 here -- so the centred bookkeeping kernels still run.
template <class T>
struct CurveFittingSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; const T* data; int nEdges; const int *gd, *gp; };
    long N, Usz; bool usePreconditioner = true, usesGraph = true;
    CurveFittingSpec(const unsigned* d) : N(d[0]), Usz(d[1]) {}
    static const char* name() { return "curveFitting"; }
    int numArrays() const { return 1; }
    void describe(int, long& el, int& ch) const { el = Usz; ch = 2; }
    void bind(void** q, Params& P) const { P.U[0].data = (T*)q[0]; P.data = (const T*)q[1]; P.nEdges = *(const int*)q[2]; P.gd = (const int*)q[3]; P.gp = (const int*)q[4]; }
    int numGroups(const Params&) const { return 2; }
    GeGroup group(const Params& P, int g) const { return g == 0 ? GeGroup{ Usz, 1, 1, 1, 0 } : GeGroup{ (long)P.nEdges, 1, 2, 0, 0 }; }
    __device__ static bool excluded(const Params&, int, long) { return false; }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        if (g == 0) { out[0] = S(T(0)); return; }
        const long d = P.gd[i], p = P.gp[i];
        const T x = P.data[2 * d], y = P.data[2 * d + 1];
        const S a = X(0, p, 0), b = X(0, p, 1);
        out[0] = y - (a * geCos(b * x) + b * geSin(a * x));
    }
};

// ============================================================ shape_from_shading.t:1-90  (LM in the benchmark config; no UsePreconditioner call)
// ComputedArrays B_I / valid (o.t:1007-1040) are evaluated inline: this IS kernels.precompute.
template <class T>
struct ShapeFromShadingSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; int W, H; T wp, ws, wg, fx, fy, ux, uy, L[10]; const T *D, *Im; const unsigned char *eR, *eC; };
    int W, H; bool usePreconditioner = false, usesGraph = false;
    ShapeFromShadingSpec(const unsigned* d) : W((int)d[0]), H((int)d[1]) {}
    static const char* name() { return "shape_from_shading"; }
    int numArrays() const { return 1; }
    void describe(int, long& el, int& ch) const { el = (long)W * H; ch = 1; }
    void bind(void** q, Params& P) const {
        P.W = W; P.H = H;
        P.wp = sqrt((T) * (const float*)q[0]); P.ws = sqrt((T) * (const float*)q[1]); P.wg = sqrt((T) * (const float*)q[2]);   // spec lines 4-6: sqrt(Param)
        P.fx = (T) * (const float*)q[3]; P.fy = (T) * (const float*)q[4]; P.ux = (T) * (const float*)q[5]; P.uy = (T) * (const float*)q[6];
        for (int i = 1; i <= 9; ++i) P.L[i] = (T) * (const float*)q[6 + i];
        P.U[0].data = (T*)q[16]; P.D = (const T*)q[17]; P.Im = (const T*)q[18]; P.eR = (const unsigned char*)q[19]; P.eC = (const unsigned char*)q[20];
    }
    int numGroups(const Params&) const { return 4; }
    GeGroup group(const Params&, int g) const { return GeGroup{ (long)W * H, g == 3 ? 3 : 1, g == 0 ? 1 : 5, 1, 0 }; }
    __device__ static bool excluded(const Params& P, int, long e) { return !(P.D[e] > T(0)); }
    __device__ static long id(const Params& P, int x, int y) { return (long)y * P.W + x; }
    __device__ static bool expanded1(const Params& P, int x, int y) { return x - 1 >= 0 && x + 1 < P.W && y - 1 >= 0 && y + 1 < P.H; }
    __device__ static bool dv(const Params& P, int x, int y) { return P.D[id(P, x, y)] > T(0); }
    template <class A> __device__ static typename A::S BI(const Params& P, A& X, int px, int py) {   // spec lines 27-66
        using S = typename A::S;
        if (px < 0 || px >= P.W || py < 0 || py >= P.H || !expanded1(P, px, py)) return S(T(0));
        if (!(dv(P, px - 1, py) && dv(P, px, py) && dv(P, px, py - 1))) return S(T(0));
        const T i = T(px), j = T(py);
        const S x00 = X(0, id(P, px, py), 0), xm0 = X(0, id(P, px - 1, py), 0), x0m = X(0, id(P, px, py - 1), 0);
        S nx = x0m * (x00 - xm0) / P.fy;
        S ny = xm0 * (x00 - x0m) / P.fx;
        S nz = (nx * (P.ux - i) / P.fx) + (ny * (P.uy - j) / P.fy) - (xm0 * x0m / (P.fx * P.fy));
        const S sq = nx * nx + ny * ny + nz * nz;
        const S inv = geVal(sq) > T(0) ? T(1) / geSqrt(sq) : S(T(1));
        nx = inv * nx; ny = inv * ny; nz = inv * nz;
        const S B = P.L[1] + P.L[2] * ny + P.L[3] * nz + P.L[4] * nx + P.L[5] * nx * ny + P.L[6] * ny * nz
                  + P.L[7] * (-(nx * nx) - ny * ny + T(2) * nz * nz) + P.L[8] * nz * nx + P.L[9] * (nx * nx - ny * ny);
        const T I = P.Im[id(P, px, py)] * T(0.5) + T(0.25) * (P.Im[id(P, px - 1, py)] + P.Im[id(P, px, py - 1)]);
        return B - I;
    }
    template <class A> __device__ static void eval(const Params& P, int g, long ii, A& X, typename A::S* out) {
        using S = typename A::S;
        const int x = (int)(ii % P.W), y = (int)(ii / P.W);
        if (g == 0) {
            const S e = P.wp * (X(0, ii, 0) - P.D[ii]);
            out[0] = dv(P, x, y) ? e : S(T(0));
        } else if (g == 1 || g == 2) {
            if (!expanded1(P, x, y)) { out[0] = S(T(0)); return; }
            const S a = BI(P, X, x, y);
            const S b = (g == 1) ? BI(P, X, x + 1, y) : BI(P, X, x, y + 1);
            const T m = (g == 1) ? T(P.eR[ii]) : T(P.eC[ii]);
            out[0] = P.wg * ((a - b) * m);
        } else {
            out[0] = out[1] = out[2] = S(T(0));
            if (!expanded1(P, x, y)) return;
            const S c = X(0, ii, 0), l = X(0, id(P, x - 1, y), 0), u = X(0, id(P, x, y - 1), 0), rr = X(0, id(P, x + 1, y), 0), d = X(0, id(P, x, y + 1), 0);
            const T thre = T(0.01);
            const bool valid = dv(P, x, y) && dv(P, x, y - 1) && dv(P, x, y + 1) && dv(P, x - 1, y) && dv(P, x + 1, y)
                && fabs(geVal(c) - geVal(u)) < thre && fabs(geVal(c) - geVal(d)) < thre && fabs(geVal(c) - geVal(l)) < thre && fabs(geVal(c) - geVal(rr)) < thre;
            if (!valid) return;
            const T i0 = T(x), j0 = T(y);
            // p(ox,oy) = (((i-u_x)/f_x) d, ((j-u_y)/f_y) d, d);  E_s = 4 p(0,0) - (p(-1,0)+p(0,-1)+p(1,0)+p(0,1))
            out[0] = P.ws * (T(4.0) * (((i0 - P.ux) / P.fx) * c) - ((((i0 - T(1) - P.ux) / P.fx) * l) + (((i0 - P.ux) / P.fx) * u) + (((i0 + T(1) - P.ux) / P.fx) * rr) + (((i0 - P.ux) / P.fx) * d)));
            out[1] = P.ws * (T(4.0) * (((j0 - P.uy) / P.fy) * c) - ((((j0 - P.uy) / P.fy) * l) + (((j0 - T(1) - P.uy) / P.fy) * u) + (((j0 - P.uy) / P.fy) * rr) + (((j0 + T(1) - P.uy) / P.fy) * d)));
            out[2] = P.ws * (T(4.0) * c - (l + u + rr + d));
        }
    }
};

// ============================================================ robust_nonrigid_alignment.t:1-28 (fp64 LM in the benchmark config)
template <class T>
struct RobustNonrigidSpec {
    struct Params { GeArray<T> U[kGeMaxArrays]; long N; int nEdges; const int *v0, *v1; const T *UrShape, *Constraints, *Normals; T wfit, wreg; };
    long N; bool usePreconditioner = true, usesGraph = true;
    RobustNonrigidSpec(const unsigned* d) : N(d[0]) {}
    static const char* name() { return "robust_nonrigid_alignment"; }
    int numArrays() const { return 3; }
    void describe(int a, long& el, int& ch) const { el = N; ch = a == 2 ? 1 : 3; }
    void bind(void** q, Params& P) const {
        P.N = N; P.wfit = (T) * (const float*)q[0]; P.wreg = (T) * (const float*)q[1];
        P.U[0].data = (T*)q[2]; P.U[1].data = (T*)q[3]; P.U[2].data = (T*)q[4];
        P.UrShape = (const T*)q[5]; P.Constraints = (const T*)q[6]; P.Normals = (const T*)q[7];
        P.nEdges = *(const int*)q[8]; P.v0 = (const int*)q[9]; P.v1 = (const int*)q[10];
    }
    int numGroups(const Params&) const { return 3; }
    GeGroup group(const Params& P, int g) const { return g == 0 ? GeGroup{ N, 3, 4, 1, 0 } : (g == 1 ? GeGroup{ N, 3, 1, 1, 0 } : GeGroup{ (long)P.nEdges, 3, 9, 0, 0 }); }
    __device__ static bool excluded(const Params&, int, long) { return false; }
    template <class A> __device__ static void eval(const Params& P, int g, long i, A& X, typename A::S* out) {
        using S = typename A::S;
        if (g == 0) {   // w_fit * Select(Constraints >= -999999.9, rw * n.(O - C), 0): Select lifts over the 3-vector condition (ad.t:330-348)
            const S rw = X(2, i, 0);
            const S dotp = P.Normals[3 * i] * (X(0, i, 0) - P.Constraints[3 * i]) + P.Normals[3 * i + 1] * (X(0, i, 1) - P.Constraints[3 * i + 1])
                         + P.Normals[3 * i + 2] * (X(0, i, 2) - P.Constraints[3 * i + 2]);
            const S e = rw * dotp;
            for (int c = 0; c < 3; ++c) out[c] = P.wfit * (P.Constraints[3 * i + c] >= T(-999999.9) ? e : S(T(0)));
        } else if (g == 1) {
            const S rw = X(2, i, 0);
            const S e = T(1) - rw * rw;
            for (int c = 0; c < 3; ++c) out[c] = T(0.1) * (P.Constraints[3 * i + c] >= T(-999999.9) ? e : S(T(0)));
        } else {
            const long a = P.v0[i], b = P.v1[i];
            const G3<S> ang = { X(1, a, 0), X(1, a, 1), X(1, a, 2) };
            const G3<S> d = { S(P.UrShape[3 * a] - P.UrShape[3 * b]), S(P.UrShape[3 * a + 1] - P.UrShape[3 * b + 1]), S(P.UrShape[3 * a + 2] - P.UrShape[3 * b + 2]) };
            const G3<S> r = geRotate3D(ang, d);
            out[0] = P.wreg * ((X(0, a, 0) - X(0, b, 0)) - r.x);
            out[1] = P.wreg * ((X(0, a, 1) - X(0, b, 1)) - r.y);
            out[2] = P.wreg * ((X(0, a, 2) - X(0, b, 2)) - r.z);
        }
    }
};

template <class T>
EnergyOps<T>* makeGenericEnergy(const std::string& energy, const unsigned* dims) {
    if (energy == "image_warping") return new GenericOps<ImageWarpingSpec<T>, T>(dims);
    if (energy == "arap_mesh_deformation") return new GenericOps<ArapSpec<T>, T>(dims);
    if (energy == "poisson_image_editing") return new GenericOps<PoissonSpec<T>, T>(dims);
    if (energy == "laplacian") return new GenericOps<LaplacianSpec<T>, T>(dims);
    if (energy == "curveFitting") return new GenericOps<CurveFittingSpec<T>, T>(dims);
    if (energy == "shape_from_shading") return new GenericOps<ShapeFromShadingSpec<T>, T>(dims);
    if (energy == "robust_nonrigid_alignment") return new GenericOps<RobustNonrigidSpec<T>, T>(dims);
    return nullptr;
}
template EnergyOps<float>* makeGenericEnergy<float>(const std::string&, const unsigned*);
template EnergyOps<double>* makeGenericEnergy<double>(const std::string&, const unsigned*);

}  // namespace optb
