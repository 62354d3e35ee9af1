// energy_arap.cu -- hand-written sm_100a kernels for arap_mesh_deformation (3-D ARAP over a vertex graph).
//
// Energy (reference examples/arap_mesh_deformation/arap_mesh_deformation.t:1-18), N vertices, directed edges
// e = (a -> b) given by the graph's v0/v1 index arrays:
//     r_fit(v) = w_fit (O_v - C_v)                         valid iff C_v.x >= -999999.9
//     r_reg(e) = w_reg [ (O_a - O_b) - R(A_a) (U_a - U_b) ]        R = Rotate3D (lib.t:77-91)
// The reference evaluates the graph term per edge and scatters nine atomicAdds per edge into Ap / r / pre
// (createjtjgraph o.t:2092-2126, createjtfgraph :2228-2253, kernels :687-706), evaluating sin/cos of the
// angles per edge per PCG step.  Here the graph is turned into two vertex-CSR neighbour lists once per
// solve (stable cub radix sort of the caller's index arrays) and every kernel is a per-vertex GATHER:
// deterministic, no atomics.  Nothing is stored per edge: PCGInit1 packs per VERTEX (sin/cos of the three
// angles, rest position, fit flag) = 48 bytes, and the pair Jacobian block M_e = d(R(A_a) d_e)/dA_a is rebuilt in
// registers from the endpoint's pack (the whole pack array, 24 MB at 500 k vertices, lives in L2).
//     (J^T J p)_O(v) = w^2 [ sum_{e=(v,b)} Jp_e  -  sum_{e=(c,v)} Jp_e ] + w_fit^2 fit_v pO_v ,  Jp_e = (pO_a - pO_b) - M_e pA_a
//     (J^T J p)_A(v) = -w^2 sum_{e=(v,b)} M_e^T Jp_e
#include <cub/cub.cuh>
#include "solver.h"

namespace optb {

template <class T> struct alignas(16) ArapPack { T sa, ca, sb, cb, sg, cg, ux, uy, uz, fit, pad0, pad1; };
template <class T> struct V3 { T x, y, z; };

template <class T>
struct ArapData {
    int N;
    const T *Offset, *Angle, *UrShape, *Constraints;   // caller arrays, 3 channels each
    T w, wf;
    long offA;                                         // Angle block inside the flat PCG vectors
    ArapPack<T>* pack;
    const int *outStart, *outNbr, *inStart, *inNbr;    // vertex-CSR neighbour lists (out: edges v->b, in: edges c->v)
    int symmetric;                                     // every edge (a,b) has its reverse (b,a): the in-list of a vertex equals its out-list
    int lo, hi, wlo, whi;                              // multi-GPU: owned vertices [lo,hi) and the window incl. halo [wlo,whi); single GPU: [0,N)
};

__device__ __forceinline__ void sincosV(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincosV(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ float guardedInvertA(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }
__device__ __forceinline__ double guardedInvertA(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }

// rotation R(A) and the three columns of M = d(R d)/dA for one rest edge d, from the packed sin/cos
template <class T>
struct ArapRot {
    T m0, m1, m2, m3, m4, m5, m6, m7, m8, sa, ca, sb, cb, sg, cg;
    __device__ __forceinline__ explicit ArapRot(const ArapPack<T>& s) : sa(s.sa), ca(s.ca), sb(s.sb), cb(s.cb), sg(s.sg), cg(s.cg) {
        m0 = cg * cb; m1 = -sg * ca + cg * sb * sa; m2 = sg * sa + cg * sb * ca;
        m3 = sg * cb; m4 = cg * ca + sg * sb * sa; m5 = -cg * sa + sg * sb * ca;
        m6 = -sb; m7 = cb * sa; m8 = cb * ca;
    }
    __device__ __forceinline__ V3<T> apply(const V3<T>& d) const { return { m0 * d.x + m1 * d.y + m2 * d.z, m3 * d.x + m4 * d.y + m5 * d.z, m6 * d.x + m7 * d.y + m8 * d.z }; }
    // columns of M for rest edge d (Rd = R d):  dR/dalpha d, dR/dbeta d, dR/dgamma d
    __device__ __forceinline__ void columns(const V3<T>& d, const V3<T>& Rd, V3<T>& ca_, V3<T>& cb_, V3<T>& cg_) const {
        ca_ = { m2 * d.y - m1 * d.z, m5 * d.y - m4 * d.z, m8 * d.y - m7 * d.z };
        const T t = -sb * d.x + cb * (sa * d.y + ca * d.z);
        cb_ = { cg * t, sg * t, -cb * d.x - sb * (sa * d.y + ca * d.z) };
        cg_ = { -Rd.y, Rd.x, T(0) };
    }
};
template <class T> __device__ __forceinline__ V3<T> ld3(const T* p, long i) { return { p[3 * i], p[3 * i + 1], p[3 * i + 2] }; }
template <class T> __device__ __forceinline__ T dot3(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// ------------------------------------------------------------------------------------ PCGInit1 (+_Graph,+_Finish)
// pass 1: per-vertex pack.  pass 2: gather J^T F and diag(J^T J), then r, pre, p, r.p  (:361-419, :687-692)
template <class T>
__global__ void __launch_bounds__(kBlock) k_arapPack(ArapData<T> D) {
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        ArapPack<T> s;
        sincosV(D.Angle[3 * v], &s.sa, &s.ca); sincosV(D.Angle[3 * v + 1], &s.sb, &s.cb); sincosV(D.Angle[3 * v + 2], &s.sg, &s.cg);
        s.ux = D.UrShape[3 * v]; s.uy = D.UrShape[3 * v + 1]; s.uz = D.UrShape[3 * v + 2];
        s.fit = D.Constraints[3 * v] >= T(-999999.9) ? T(1) : T(0); s.pad0 = s.pad1 = T(0);
        D.pack[v] = s;
    }
}
template <class T>
__global__ void __launch_bounds__(kBlock) k_arapInit(ArapData<T> D, ReduceWorkspace ws, int rzSlot, int usePre, T* __restrict__ delta, T* __restrict__ r,
                                                     T* __restrict__ pre, T* __restrict__ p) {
    T acc = T(0);
    const T w = D.w, w2 = D.w * D.w;
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        if (v < D.lo || v >= D.hi) {   // halo vertex: no unknowns of ours; its p arrives from the owner (k_haloExchange)
            const long o = 3L * v, a = D.offA + 3L * v;
            delta[o] = delta[o + 1] = delta[o + 2] = T(0); delta[a] = delta[a + 1] = delta[a + 2] = T(0);
            r[o] = r[o + 1] = r[o + 2] = T(0); r[a] = r[a + 1] = r[a + 2] = T(0);
            pre[o] = pre[o + 1] = pre[o + 2] = T(0); pre[a] = pre[a + 1] = pre[a + 2] = T(0);
            continue;
        }
        const ArapPack<T> sv = D.pack[v];
        const ArapRot<T> Rv(sv);
        const V3<T> Ov = ld3(D.Offset, v);
        V3<T> gO{ T(0), T(0), T(0) }, gA{ T(0), T(0), T(0) }, dA{ T(0), T(0), T(0) };
        T dO = T(0);
        for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {      // edges v -> b: O_v enters with +w, A_v with -w M
            const int b = D.outNbr[k];
            const ArapPack<T> sb_ = D.pack[b];
            const V3<T> Ob = ld3(D.Offset, b);
            const V3<T> d = { sv.ux - sb_.ux, sv.uy - sb_.uy, sv.uz - sb_.uz };
            const V3<T> Rd = Rv.apply(d);
            const V3<T> re = { w * ((Ov.x - Ob.x) - Rd.x), w * ((Ov.y - Ob.y) - Rd.y), w * ((Ov.z - Ob.z) - Rd.z) };
            V3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
            gO.x += w * re.x; gO.y += w * re.y; gO.z += w * re.z;
            gA.x -= w * dot3(c0, re); gA.y -= w * dot3(c1, re); gA.z -= w * dot3(c2, re);
            dO += w2;
            dA.x += w2 * dot3(c0, c0); dA.y += w2 * dot3(c1, c1); dA.z += w2 * dot3(c2, c2);
        }
        for (int k = D.inStart[v]; k < D.inStart[v + 1]; ++k) {        // edges c -> v: O_v enters with -w
            const int c = D.inNbr[k];
            const ArapPack<T> sc = D.pack[c];
            const ArapRot<T> Rc(sc);
            const V3<T> Oc = ld3(D.Offset, c);
            const V3<T> d = { sc.ux - sv.ux, sc.uy - sv.uy, sc.uz - sv.uz };
            const V3<T> Rd = Rc.apply(d);
            gO.x -= w * (w * ((Oc.x - Ov.x) - Rd.x)); gO.y -= w * (w * ((Oc.y - Ov.y) - Rd.y)); gO.z -= w * (w * ((Oc.z - Ov.z) - Rd.z));
            dO += w2;
        }
        if (sv.fit != T(0)) {
            const V3<T> C = ld3(D.Constraints, v);
            gO.x += D.wf * (D.wf * (Ov.x - C.x)); gO.y += D.wf * (D.wf * (Ov.y - C.y)); gO.z += D.wf * (D.wf * (Ov.z - C.z));
            dO += D.wf * D.wf;
        }
        // graph problems: PCGInit1_Finish order -- invert first, then the UsePreconditioner(false) override (:406-410)
        T iO = guardedInvertA(dO), iAx = guardedInvertA(dA.x), iAy = guardedInvertA(dA.y), iAz = guardedInvertA(dA.z);
        if (!usePre) iO = iAx = iAy = iAz = T(1);
        const V3<T> rO = { -gO.x, -gO.y, -gO.z }, rA = { -gA.x, -gA.y, -gA.z };
        const V3<T> pO = { iO * rO.x, iO * rO.y, iO * rO.z }, pA = { iAx * rA.x, iAy * rA.y, iAz * rA.z };
        acc += dot3(rO, pO) + dot3(rA, pA);
        const long o = 3L * v, a = D.offA + 3L * v;
        delta[o] = delta[o + 1] = delta[o + 2] = T(0); delta[a] = delta[a + 1] = delta[a + 2] = T(0);
        r[o] = rO.x; r[o + 1] = rO.y; r[o + 2] = rO.z; r[a] = rA.x; r[a + 1] = rA.y; r[a + 2] = rA.z;
        pre[o] = pre[o + 1] = pre[o + 2] = iO; pre[a] = iAx; pre[a + 1] = iAy; pre[a + 2] = iAz;
        p[o] = pO.x; p[o + 1] = pO.y; p[o + 2] = pO.z; p[a] = pA.x; p[a + 1] = pA.y; p[a + 2] = pA.z;
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}

// ------------------------------------------------------------------------------------ PCGStep1 (+_Graph), optionally fused with PCGStep3
// out = (J^T J [+CtC]) p, dot = p.out  (:421-434, :694-706).  FUSE3: p = z + beta pOld formed on the fly for the vertex
// and every neighbour it reads; the vertex's own value goes to pNew (double-buffered).
template <class T, bool FUSE3, bool LM>
__global__ void __launch_bounds__(kBlock) k_arapApplyJTJ(ArapData<T> D, ReduceWorkspace ws, int dotSlot, int rzNew, int rzOld, const T* __restrict__ pIn,
                                                         const T* __restrict__ z, T* __restrict__ pOut, T* __restrict__ out, const T* __restrict__ CtC) {
    T beta = T(0);
    if (FUSE3) {
        double drn, dro; loadScalars2(ws, rzNew, rzOld, drn, dro);
        const T rn = (T)drn, ro = (T)dro;
        if (ro > T(0)) beta = rn / ro;
    }
    auto dirO = [&](int v) -> V3<T> { V3<T> q = ld3(pIn, v); if (FUSE3) { const V3<T> zz = ld3(z, v); q.x = zz.x + beta * q.x; q.y = zz.y + beta * q.y; q.z = zz.z + beta * q.z; } return q; };
    auto dirA = [&](int v) -> V3<T> { V3<T> q = ld3(pIn + D.offA, v); if (FUSE3) { const V3<T> zz = ld3(z + D.offA, v); q.x = zz.x + beta * q.x; q.y = zz.y + beta * q.y; q.z = zz.z + beta * q.z; } return q; };
    const T w2 = D.w * D.w, wf2 = D.wf * D.wf;
    T acc = T(0);
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        const V3<T> pOv = dirO(v), pAv = dirA(v);
        if (v < D.lo || v >= D.hi) {   // halo vertex: only its new direction is needed here
            const long o = 3L * v, a = D.offA + 3L * v;
            out[o] = out[o + 1] = out[o + 2] = T(0); out[a] = out[a + 1] = out[a + 2] = T(0);
            if (FUSE3) { pOut[o] = pOv.x; pOut[o + 1] = pOv.y; pOut[o + 2] = pOv.z; pOut[a] = pAv.x; pOut[a + 1] = pAv.y; pOut[a + 2] = pAv.z; }
            continue;
        }
        const ArapPack<T> sv = D.pack[v];
        const ArapRot<T> Rv(sv);
        V3<T> aO{ T(0), T(0), T(0) }, aA{ T(0), T(0), T(0) };
        for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {
            const int b = D.outNbr[k];
            const ArapPack<T> sb_ = D.pack[b];
            const V3<T> pOb = dirO(b);
            const V3<T> d = { sv.ux - sb_.ux, sv.uy - sb_.uy, sv.uz - sb_.uz };
            const V3<T> Rd = Rv.apply(d);
            V3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
            const V3<T> Jp = { (pOv.x - pOb.x) - (c0.x * pAv.x + c1.x * pAv.y + c2.x * pAv.z),
                               (pOv.y - pOb.y) - (c0.y * pAv.x + c1.y * pAv.y + c2.y * pAv.z),
                               (pOv.z - pOb.z) - (c0.z * pAv.x + c1.z * pAv.y + c2.z * pAv.z) };
            aO.x += Jp.x; aO.y += Jp.y; aO.z += Jp.z;
            aA.x -= dot3(c0, Jp); aA.y -= dot3(c1, Jp); aA.z -= dot3(c2, Jp);
            if (D.symmetric) {   // the reverse edge (b -> v) exists: account for it with the same neighbour data (one gather instead of two)
                const ArapRot<T> Rb(sb_);
                const V3<T> pAb = dirA(b);
                const V3<T> dr = { -d.x, -d.y, -d.z };                       // U_b - U_v
                const V3<T> Rdr = Rb.apply(dr);
                V3<T> e0, e1, e2; Rb.columns(dr, Rdr, e0, e1, e2);
                aO.x -= (pOb.x - pOv.x) - (e0.x * pAb.x + e1.x * pAb.y + e2.x * pAb.z);
                aO.y -= (pOb.y - pOv.y) - (e0.y * pAb.x + e1.y * pAb.y + e2.y * pAb.z);
                aO.z -= (pOb.z - pOv.z) - (e0.z * pAb.x + e1.z * pAb.y + e2.z * pAb.z);
            }
        }
        for (int k = D.symmetric ? D.inStart[v + 1] : D.inStart[v]; k < D.inStart[v + 1]; ++k) {
            const int c = D.inNbr[k];
            const ArapPack<T> sc = D.pack[c];
            const ArapRot<T> Rc(sc);
            const V3<T> pOc = dirO(c), pAc = dirA(c);
            const V3<T> d = { sc.ux - sv.ux, sc.uy - sv.uy, sc.uz - sv.uz };
            const V3<T> Rd = Rc.apply(d);
            V3<T> c0, c1, c2; Rc.columns(d, Rd, c0, c1, c2);
            aO.x -= (pOc.x - pOv.x) - (c0.x * pAc.x + c1.x * pAc.y + c2.x * pAc.z);
            aO.y -= (pOc.y - pOv.y) - (c0.y * pAc.x + c1.y * pAc.y + c2.y * pAc.z);
            aO.z -= (pOc.z - pOv.z) - (c0.z * pAc.x + c1.z * pAc.y + c2.z * pAc.z);
        }
        const T f = sv.fit != T(0) ? wf2 : T(0);
        V3<T> oO = { w2 * aO.x + f * pOv.x, w2 * aO.y + f * pOv.y, w2 * aO.z + f * pOv.z };
        V3<T> oA = { w2 * aA.x, w2 * aA.y, w2 * aA.z };
        const long o = 3L * v, a = D.offA + 3L * v;
        if (LM) {
            oO.x += CtC[o] * pOv.x; oO.y += CtC[o + 1] * pOv.y; oO.z += CtC[o + 2] * pOv.z;
            oA.x += CtC[a] * pAv.x; oA.y += CtC[a + 1] * pAv.y; oA.z += CtC[a + 2] * pAv.z;
        }
        acc += dot3(pOv, oO) + dot3(pAv, oA);
        out[o] = oO.x; out[o + 1] = oO.y; out[o + 2] = oO.z; out[a] = oA.x; out[a + 1] = oA.y; out[a + 2] = oA.z;
        if (FUSE3) { pOut[o] = pOv.x; pOut[o + 1] = pOv.y; pOut[o + 2] = pOv.z; pOut[a] = pAv.x; pOut[a + 1] = pAv.y; pOut[a + 2] = pAv.z; }
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}

// ------------------------------------------------------------------------------------ computeCost (+_Graph) / computeModelCost (+_Graph)
template <class T, bool MODEL>
__global__ void __launch_bounds__(kBlock) k_arapCost(ArapData<T> D, ReduceWorkspace ws, int slot, const T* __restrict__ delta) {
    T acc = T(0);
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        if (v < D.lo || v >= D.hi) continue;
        const V3<T> Ov = ld3(D.Offset, v), Uv = ld3(D.UrShape, v);
        ArapPack<T> sv; sincosV(D.Angle[3 * v], &sv.sa, &sv.ca); sincosV(D.Angle[3 * v + 1], &sv.sb, &sv.cb); sincosV(D.Angle[3 * v + 2], &sv.sg, &sv.cg);
        const ArapRot<T> Rv(sv);
        V3<T> dOv{ T(0), T(0), T(0) }, dAv{ T(0), T(0), T(0) };
        if (MODEL) { dOv = ld3(delta, v); dAv = ld3(delta + D.offA, v); }
        T e = T(0);
        for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {
            const int b = D.outNbr[k];
            const V3<T> Ob = ld3(D.Offset, b), Ub = ld3(D.UrShape, b);
            const V3<T> d = { Uv.x - Ub.x, Uv.y - Ub.y, Uv.z - Ub.z };
            const V3<T> Rd = Rv.apply(d);
            V3<T> re = { D.w * ((Ov.x - Ob.x) - Rd.x), D.w * ((Ov.y - Ob.y) - Rd.y), D.w * ((Ov.z - Ob.z) - Rd.z) };
            if (MODEL) {
                const V3<T> dOb = ld3(delta, b);
                V3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
                re.x += D.w * ((dOv.x - dOb.x) - (c0.x * dAv.x + c1.x * dAv.y + c2.x * dAv.z));
                re.y += D.w * ((dOv.y - dOb.y) - (c0.y * dAv.x + c1.y * dAv.y + c2.y * dAv.z));
                re.z += D.w * ((dOv.z - dOb.z) - (c0.z * dAv.x + c1.z * dAv.y + c2.z * dAv.z));
            }
            e += dot3(re, re);
        }
        if (D.Constraints[3 * v] >= T(-999999.9)) {
            const V3<T> C = ld3(D.Constraints, v);
            V3<T> f = { D.wf * (Ov.x - C.x), D.wf * (Ov.y - C.y), D.wf * (Ov.z - C.z) };
            if (MODEL) { f.x += D.wf * dOv.x; f.y += D.wf * dOv.y; f.z += D.wf * dOv.z; }
            e += dot3(f, f);
        }
        acc += T(0.5) * e;
    }
    gridReduceTo<kBlock>((double)acc, ws, slot);
}
// PCGComputeCtC (+_Graph): diag(J^T J)/radius from the pack
template <class T>
__global__ void __launch_bounds__(kBlock) k_arapCtC(ArapData<T> D, T* __restrict__ CtC, T invRadius) {
    const T w2 = D.w * D.w;
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        const ArapPack<T> sv = D.pack[v];
        const ArapRot<T> Rv(sv);
        T dO = T(0); V3<T> dA{ T(0), T(0), T(0) };
        for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {
            const ArapPack<T> sb_ = D.pack[D.outNbr[k]];
            const V3<T> d = { sv.ux - sb_.ux, sv.uy - sb_.uy, sv.uz - sb_.uz };
            const V3<T> Rd = Rv.apply(d);
            V3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
            dO += w2 * invRadius; dA.x += w2 * dot3(c0, c0) * invRadius; dA.y += w2 * dot3(c1, c1) * invRadius; dA.z += w2 * dot3(c2, c2) * invRadius;
        }
        dO += (T)(D.inStart[v + 1] - D.inStart[v]) * w2 * invRadius;
        if (sv.fit != T(0)) dO += D.wf * D.wf * invRadius;
        const long o = 3L * v, a = D.offA + 3L * v;
        CtC[o] = CtC[o + 1] = CtC[o + 2] = dO; CtC[a] = dA.x; CtC[a + 1] = dA.y; CtC[a + 2] = dA.z;
    }
}
// graph symmetry test: every out-neighbour b of v must list v among ITS out-neighbours, and in/out degrees must match
__global__ void k_arapSymmetric(int N, const int* __restrict__ outStart, const int* __restrict__ outNbr, const int* __restrict__ inStart, int* __restrict__ notSymmetric) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
        bool ok = (outStart[v + 1] - outStart[v]) == (inStart[v + 1] - inStart[v]);
        for (int k = outStart[v]; ok && k < outStart[v + 1]; ++k) {
            const int b = outNbr[k];
            int cnt = 0, mine = 0;
            for (int q = outStart[b]; q < outStart[b + 1]; ++q) cnt += outNbr[q] == v;
            for (int q = outStart[v]; q < outStart[v + 1]; ++q) mine += outNbr[q] == b;
            ok = cnt == mine;                                            // multiplicities must match too (multigraphs)
        }
        if (!ok) *notSymmetric = 1;
    }
}
__global__ void k_csrStart(int nE, const int* __restrict__ sortedKeys, int N, int* __restrict__ start) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= nE; t += gridDim.x * blockDim.x) {
        const int k = t < nE ? sortedKeys[t] : N, kp = t > 0 ? sortedKeys[t - 1] : -1;
        for (int u = kp + 1; u <= k && u <= N; ++u) start[u] = t;
    }
}

// ====================================================================================== host side
template <class T>
struct ArapOps : EnergyOps<T> {
    ArapData<T> D{};
    const int *v0 = nullptr, *v1 = nullptr; int nEdges = 0;
    int *outStart = nullptr, *outNbr = nullptr, *inStart = nullptr, *inNbr = nullptr, *keysTmp = nullptr;
    void* cubTemp = nullptr; size_t cubBytes = 0; int capE = 0;
    bool csrValid = false;
    int sms = 148;
    ArapOps(const unsigned* dims) {
        D.N = (int)dims[0]; D.lo = D.wlo = 0; D.hi = D.whi = D.N;
        this->unknowns.resize(2);
        this->unknowns[0].elements = D.N; this->unknowns[0].channels = 3;   // Offset
        this->unknowns[1].elements = D.N; this->unknowns[1].channels = 3;   // Angle
        this->layoutUnknowns();
        D.offA = this->unknowns[1].offset;
        this->usePreconditioner = true; this->usesGraph = true;
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        OPT_CUDA_CHECK(cudaMalloc((void**)&D.pack, sizeof(ArapPack<T>) * (size_t)D.N));
        OPT_CUDA_CHECK(cudaMalloc((void**)&outStart, sizeof(int) * (D.N + 2))); OPT_CUDA_CHECK(cudaMalloc((void**)&inStart, sizeof(int) * (D.N + 2)));
    }
    ~ArapOps() override { for (void* q : { (void*)D.pack, (void*)outStart, (void*)outNbr, (void*)inStart, (void*)inNbr, (void*)keysTmp, cubTemp }) if (q) cudaFree(q); }
    const char* name() const override { return "arap_mesh_deformation"; }
    bool capturable() const override { return csrValid; }   // the neighbour lists are built (with a read-back) at the first use after Opt_ProblemInit
    unsigned long long captureKey() const override { return EnergyOps<T>::hashValue(D); }
    int grid() const { long g = ((long)(D.whi - D.wlo) + kBlock - 1) / kBlock; long cap = (long)sms * 8; if (g > cap) g = cap; if (g > kMaxPartials) g = kMaxPartials; return g < 1 ? 1 : (int)g; }
    void bind(void** q) override {   // positions: arap_mesh_deformation.t:2-8
        D.wf = (T) * (const float*)q[0]; D.w = (T) * (const float*)q[1];
        this->unknowns[0].user = q[2]; this->unknowns[1].user = q[3];
        D.Offset = (const T*)q[2]; D.Angle = (const T*)q[3]; D.UrShape = (const T*)q[4]; D.Constraints = (const T*)q[5];
        const int n = *(const int*)q[6];
        if (n != nEdges || v0 != (const int*)q[7] || v1 != (const int*)q[8]) csrValid = false;
        nEdges = n; v0 = (const int*)q[7]; v1 = (const int*)q[8];
    }
    void newSolve() override { csrValid = false; }
    // ---- multi-GPU: 1-D decomposition of the (locality-ordered) vertex index.  Every rank binds the GLOBAL arrays; it owns the
    // vertices [ownedBegin, ownedEnd) and keeps `up` / `dn` halo vertices beyond them (the graph's index bandwidth, uniform).
    int haloUp = 0, haloDn = 0;
    bool setSlab(int ownedBegin, int ownedEnd, int up, int dn) override {
        if (ownedBegin < 0 || ownedEnd > D.N || ownedBegin >= ownedEnd || ownedBegin % 4 || (ownedEnd % 4 && ownedEnd != D.N) || up % 4 || dn % 4) return false;
        if (ownedBegin - up < 0 || ownedEnd + dn > D.N || up > ownedEnd - ownedBegin || dn > ownedEnd - ownedBegin) return false;
        D.lo = ownedBegin; D.hi = ownedEnd; D.wlo = ownedBegin - up; D.whi = ownedEnd + dn; haloUp = up; haloDn = dn;
        return true;
    }
    std::vector<typename EnergyOps<T>::HaloSegment> haloSegments() const override {
        std::vector<typename EnergyOps<T>::HaloSegment> s;   // vectors are global-size on every rank: destination offsets equal source offsets
        if (haloUp > 0) { s.push_back({ 3L * D.lo, 3L * D.lo, 3L * haloUp, 0 }); s.push_back({ D.offA + 3L * D.lo, D.offA + 3L * D.lo, 3L * haloUp, 0 }); }
        if (haloDn > 0) { s.push_back({ 3L * (D.hi - haloDn), 3L * (D.hi - haloDn), 3L * haloDn, 1 }); s.push_back({ D.offA + 3L * (D.hi - haloDn), D.offA + 3L * (D.hi - haloDn), 3L * haloDn, 1 }); }
        return s;
    }
    std::vector<std::pair<long, long>> ownedScalarRanges() const override { return { { 3L * D.lo, 3L * D.hi }, { D.offA + 3L * D.lo, D.offA + 3L * D.hi } }; }
    std::vector<std::pair<long, long>> processScalarRanges() const override {
        if (D.wlo == 0 && D.whi == D.N) return {};
        return { { 3L * D.wlo, 3L * D.whi }, { D.offA + 3L * D.wlo, D.offA + 3L * D.whi } };
    }   // the caller may have rewritten the index arrays in place
    void buildCsr(cudaStream_t st) {
        if (csrValid) return;
        if (nEdges > capE) {
            for (void* q : { (void*)outNbr, (void*)inNbr, (void*)keysTmp, cubTemp }) if (q) cudaFree(q);
            capE = nEdges;
            OPT_CUDA_CHECK(cudaMalloc((void**)&outNbr, sizeof(int) * capE)); OPT_CUDA_CHECK(cudaMalloc((void**)&inNbr, sizeof(int) * capE)); OPT_CUDA_CHECK(cudaMalloc((void**)&keysTmp, sizeof(int) * capE));
            cubBytes = 0; cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, v0, keysTmp, v1, outNbr, capE);
            OPT_CUDA_CHECK(cudaMalloc(&cubTemp, cubBytes));
        }
        int endBit = 1; while ((1L << endBit) <= D.N && endBit < 31) ++endBit;
        const int g = std::max(1, std::min(sms * 8, (nEdges + 256) / 256));
        cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, v0, keysTmp, v1, outNbr, nEdges, 0, endBit, st);   // stable: neighbours keep the caller's edge order
        k_csrStart<<<g, 256, 0, st>>>(nEdges, keysTmp, D.N, outStart);
        cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, v1, keysTmp, v0, inNbr, nEdges, 0, endBit, st);
        k_csrStart<<<g, 256, 0, st>>>(nEdges, keysTmp, D.N, inStart);
        D.outStart = outStart; D.outNbr = outNbr; D.inStart = inStart; D.inNbr = inNbr;
        // symmetric meshes (the reference app builds both directions of every mesh edge, OptGraph.h:64-75) take the merged loop
        int* flag = keysTmp;   // reuse: the sort keys are no longer needed
        OPT_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
        k_arapSymmetric<<<g, 256, 0, st>>>(D.N, outStart, outNbr, inStart, flag);
        int ns = 1; OPT_CUDA_CHECK(cudaMemcpyAsync(&ns, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); OPT_CUDA_CHECK(cudaStreamSynchronize(st));
        D.symmetric = (ns == 0 && !getenv("OPTB_ARAP_NO_SYMMETRIC")) ? 1 : 0;
        csrValid = true;
    }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        buildCsr(c.stream);
        k_arapPack<T><<<grid(), kBlock, 0, c.stream>>>(D);
        k_arapInit<T><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, rzSlot, this->usePreconditioner ? 1 : 0, delta, r, pre, p);
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan*) override {
        if (CtC) k_arapApplyJTJ<T, false, true><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, 0, 0, x, nullptr, nullptr, out, CtC);
        else k_arapApplyJTJ<T, false, false><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, 0, 0, x, nullptr, nullptr, out, nullptr);
    }
    bool step3ApplyJTJ(const LaunchCtx& c, const T* z, const T* pOld, T* pNew, T* out, const T* CtC, int rzNew, int rzOld, int dotSlot, const HaloPlan*) override {
        // measured (profiles/r1_other_configs.md): fusing PCGStep3 into this gather doubles the per-neighbour reads (z and p) and is
        // slower than a separate streaming Step3 for this energy; kept as an opt-in
        if (!getenv("OPTB_ARAP_FUSION")) return false;
        if (dotSlot == -2) return true;
        if (CtC) k_arapApplyJTJ<T, true, true><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, rzNew, rzOld, pOld, z, pNew, out, CtC);
        else k_arapApplyJTJ<T, true, false><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, rzNew, rzOld, pOld, z, pNew, out, nullptr);
        return true;
    }
    void cost(const LaunchCtx& c) override { buildCsr(c.stream); k_arapCost<T, false><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, kCost, nullptr); }
    void modelCost(const LaunchCtx& c, const T* delta) override { k_arapCost<T, true><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, kModelCost, delta); }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override { k_arapCtC<T><<<grid(), kBlock, 0, c.stream>>>(D, CtC, invRadius); }
};

template <class T> EnergyOps<T>* makeArapMesh(const unsigned* dims) { return new ArapOps<T>(dims); }
template EnergyOps<float>* makeArapMesh<float>(const unsigned*);
template EnergyOps<double>* makeArapMesh<double>(const unsigned*);

}  // namespace optb
