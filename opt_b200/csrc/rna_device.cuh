// rna_device.cuh -- per-vertex math of the hand-written robust_nonrigid_alignment kernel set (energy_rna.cu).
//
// Energy (reference examples/robust_nonrigid_alignment/robust_nonrigid_alignment.t:1-28; specs/ here), N vertices with
// unknowns Offset O (3), Angle A (3), RobustWeights g (1) and directed edges e = (a -> b) from the graph's index arrays:
//     r_plane,c(v) = w_fit g_v n_v.(O_v - C_v)               if C_v,c >= -999999.9   (point-to-plane, scaled by the robust weight)
//     r_conf,c(v)  = 0.1 (1 - g_v^2)                         if C_v,c >= -999999.9
//   c = x, y, z: the spec's Select() tests the 3-vector Constraints(0) channel by channel, so both rows exist once per valid
//   channel (robust_nonrigid_alignment.t:17-19): m_v identical copies, m_v = number of valid channels (0 or 3
//   in practice)
//     r_reg(e)   = w_reg [ (O_a - O_b) - R(A_a) (U_a - U_b) ]       R = Rotate3D (lib.t:77-91)
// UsePreconditioner(true), no Exclude.  The reference scatters the graph term per edge with atomicAdds (o.t:2092-2126,
// :2228-2253, solverGPUGaussNewton.t:687-706); here, as in energy_arap.cu, the graph becomes two vertex-CSR neighbour lists
// and everything is a per-vertex GATHER (deterministic, no atomics):
//     (J^T J p)_O(v) = w^2 [ sum_{e=(v,b)} Jp_e - sum_{e=(c,v)} Jp_e ] + a1O (a1O.pO_v + a1g pg_v)     Jp_e = (pO_a - pO_b) - M_e pA_a
//     (J^T J p)_A(v) = -w^2 sum_{e=(v,b)} M_e^T Jp_e
//     (J^T J p)_g(v) = a1g (a1O.pO_v + a1g pg_v) + a2g^2 pg_v
//   with a1O = w_fit g_v n_v,  a1g = w_fit n_v.(O_v - C_v),  a2g = -0.2 g_v  (vertex terms times m_v)  and M_e = d(R(A_a) d_e)/dA_a.
// Once per linearisation a 16-scalar pack per vertex holds sin/cos of the angles, the rest position, g, n, n.(O-C) and the
// target flag; nothing is stored per edge.
//
// Everything here is plain per-vertex code callable from host and device: tests/host/rna_host.cpp runs the same functions on
// the CPU against the oracle (cost, J^T F, diag J^T J, J^T J x, model cost), so the kernels in energy_rna.cu only add the
// one-thread-per-vertex loops, the CSR construction (as in energy_arap.cu) and the reductions.
#pragma once
#if defined(__CUDACC__)
#define RNA_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define RNA_HD inline
#endif

namespace optb {

template <class T> struct alignas(16) RnaPack { T sa, ca, sb, cb, sg, cg, ux, uy, uz, g, nx, ny, nz, s, tgt, pad; };   // 16-byte aligned: the gathers load it with 128-bit loads (they are bound by L1 requests, not bytes)   // s = n.(O - C); tgt = m_v
template <class T> struct RnaV3 { T x, y, z; };

template <class T>
struct RnaData {
    int N;
    const T *Offset, *Angle, *Weights, *UrShape, *Constraints, *Normals;   // caller arrays (3,3,1,3,3,3 channels)
    T w, wf, wc;                 // w_regSqrt, w_fitSqrt, w_confSqrt (= 0.1, a constant of the spec)
    long offA, offG;             // Angle / RobustWeights blocks inside the flat PCG vectors (Offset block starts at 0)
    RnaPack<T>* pack;
    const int *outStart, *outNbr, *inStart, *inNbr;   // vertex-CSR neighbour lists (out: edges v->b, in: edges c->v)
    int lo, hi, wlo, whi;        // multi-GPU (kernels only): owned vertices [lo,hi) and the window incl. halo [wlo,whi); single GPU: [0,N)
    int symmetric;               // every edge (a,b) has its reverse (b,a) with the same multiplicity: the in-list of a vertex equals its out-list
};

RNA_HD void rnaSinCos(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
RNA_HD void rnaSinCos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
template <class T> RNA_HD RnaV3<T> rnaLd3(const T* p, long i) { return RnaV3<T>{ p[3 * i], p[3 * i + 1], p[3 * i + 2] }; }
template <class T> RNA_HD T rnaDot(const RnaV3<T>& a, const RnaV3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// rotation R(A) (lib.t:77-91) and the three columns of M = d(R d)/dA for one rest edge d, from packed sin/cos
template <class T>
struct RnaRot {
    T m0, m1, m2, m3, m4, m5, m6, m7, m8, sa, ca, sb, cb, sg, cg;
    RNA_HD explicit RnaRot(const RnaPack<T>& s) : sa(s.sa), ca(s.ca), sb(s.sb), cb(s.cb), sg(s.sg), cg(s.cg) {
        m0 = cg * cb; m1 = -sg * ca + cg * sb * sa; m2 = sg * sa + cg * sb * ca;
        m3 = sg * cb; m4 = cg * ca + sg * sb * sa; m5 = -cg * sa + sg * sb * ca;
        m6 = -sb; m7 = cb * sa; m8 = cb * ca;
    }
    RNA_HD RnaV3<T> apply(const RnaV3<T>& d) const { return RnaV3<T>{ m0 * d.x + m1 * d.y + m2 * d.z, m3 * d.x + m4 * d.y + m5 * d.z, m6 * d.x + m7 * d.y + m8 * d.z }; }
    RNA_HD void columns(const RnaV3<T>& d, const RnaV3<T>& Rd, RnaV3<T>& c0, RnaV3<T>& c1, RnaV3<T>& c2) const {   // dR/dalpha d, dR/dbeta d, dR/dgamma d
        c0 = RnaV3<T>{ m2 * d.y - m1 * d.z, m5 * d.y - m4 * d.z, m8 * d.y - m7 * d.z };
        const T t = -sb * d.x + cb * (sa * d.y + ca * d.z);
        c1 = RnaV3<T>{ cg * t, sg * t, -cb * d.x - sb * (sa * d.y + ca * d.z) };
        c2 = RnaV3<T>{ -Rd.y, Rd.x, T(0) };
    }
};

// per-linearisation pack of vertex v
template <class T> RNA_HD RnaPack<T> rnaMakePack(const RnaData<T>& D, int v) {
    RnaPack<T> s;
    rnaSinCos(D.Angle[3 * v], &s.sa, &s.ca); rnaSinCos(D.Angle[3 * v + 1], &s.sb, &s.cb); rnaSinCos(D.Angle[3 * v + 2], &s.sg, &s.cg);
    s.ux = D.UrShape[3 * v]; s.uy = D.UrShape[3 * v + 1]; s.uz = D.UrShape[3 * v + 2];
    s.g = D.Weights[v];
    s.nx = D.Normals[3 * v]; s.ny = D.Normals[3 * v + 1]; s.nz = D.Normals[3 * v + 2];
    int m = 0;
    for (int c = 0; c < 3; ++c) m += D.Constraints[3 * v + c] >= T(-999999.9) ? 1 : 0;
    const bool tgt = m > 0;
    s.tgt = T(m);
    s.s = tgt ? (s.nx * (D.Offset[3 * v] - D.Constraints[3 * v]) + s.ny * (D.Offset[3 * v + 1] - D.Constraints[3 * v + 1]) + s.nz * (D.Offset[3 * v + 2] - D.Constraints[3 * v + 2])) : T(0);
    s.pad = T(0);
    return s;
}

template <class T> struct RnaOut { RnaV3<T> O, A; T g; };

// Gather for vertex v.  MODE 0: (J^T J vec)(v)   MODE 1: (J^T F)(v)   MODE 2: diag(J^T J)(v).   vec = flat PCG vector (MODE 0)
template <class T, int MODE> RNA_HD RnaOut<T> rnaGather(const RnaData<T>& D, int v, const T* vec) {
    const RnaPack<T> sv = D.pack[v];
    const RnaRot<T> Rv(sv);
    const T w = D.w, w2 = D.w * D.w;
    RnaV3<T> aO{ T(0), T(0), T(0) }, aA{ T(0), T(0), T(0) };
    RnaV3<T> pOv{ T(0), T(0), T(0) }, pAv{ T(0), T(0), T(0) }, Ov{ T(0), T(0), T(0) };
    T pgv = T(0);
    if (MODE == 0) { pOv = rnaLd3(vec, v); pAv = rnaLd3(vec + D.offA, v); pgv = vec[D.offG + v]; }
    if (MODE == 1) Ov = rnaLd3(D.Offset, v);
    for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {          // edges v -> b
        const int b = D.outNbr[k];
        const RnaPack<T> sb_ = D.pack[b];
        const RnaV3<T> d{ sv.ux - sb_.ux, sv.uy - sb_.uy, sv.uz - sb_.uz };
        const RnaV3<T> Rd = Rv.apply(d);
        RnaV3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
        if (MODE == 2) {
            aO.x += T(1); aA.x += rnaDot(c0, c0); aA.y += rnaDot(c1, c1); aA.z += rnaDot(c2, c2);
        } else {
            RnaV3<T> y;   // row . vec / w   (MODE 0)   or   residual / w   (MODE 1)
            if (MODE == 0) {
                const RnaV3<T> pOb = rnaLd3(vec, b);
                y = RnaV3<T>{ (pOv.x - pOb.x) - (c0.x * pAv.x + c1.x * pAv.y + c2.x * pAv.z),
                              (pOv.y - pOb.y) - (c0.y * pAv.x + c1.y * pAv.y + c2.y * pAv.z),
                              (pOv.z - pOb.z) - (c0.z * pAv.x + c1.z * pAv.y + c2.z * pAv.z) };
            } else {
                const RnaV3<T> Ob = rnaLd3(D.Offset, b);
                y = RnaV3<T>{ (Ov.x - Ob.x) - Rd.x, (Ov.y - Ob.y) - Rd.y, (Ov.z - Ob.z) - Rd.z };
            }
            aO.x += y.x; aO.y += y.y; aO.z += y.z;
            aA.x -= rnaDot(c0, y); aA.y -= rnaDot(c1, y); aA.z -= rnaDot(c2, y);
            if (MODE == 0 && D.symmetric) {   // the reverse edge b -> v exists: its term needs the same neighbour data (one gather instead of two)
                const RnaRot<T> Rb(sb_);
                const RnaV3<T> pAb = rnaLd3(vec + D.offA, b), pOb = rnaLd3(vec, b);
                const RnaV3<T> dr{ -d.x, -d.y, -d.z };                       // U_b - U_v
                const RnaV3<T> Rdr = Rb.apply(dr);
                RnaV3<T> e0, e1, e2; Rb.columns(dr, Rdr, e0, e1, e2);
                aO.x -= (pOb.x - pOv.x) - (e0.x * pAb.x + e1.x * pAb.y + e2.x * pAb.z);
                aO.y -= (pOb.y - pOv.y) - (e0.y * pAb.x + e1.y * pAb.y + e2.y * pAb.z);
                aO.z -= (pOb.z - pOv.z) - (e0.z * pAb.x + e1.z * pAb.y + e2.z * pAb.z);
            }
        }
    }
    for (int k = (MODE == 0 && D.symmetric) ? D.inStart[v + 1] : D.inStart[v]; k < D.inStart[v + 1]; ++k) {            // edges c -> v: O_v enters with the opposite sign
        if (MODE == 2) { aO.x += T(1); continue; }
        const int c = D.inNbr[k];
        const RnaPack<T> sc = D.pack[c];
        const RnaRot<T> Rc(sc);
        const RnaV3<T> d{ sc.ux - sv.ux, sc.uy - sv.uy, sc.uz - sv.uz };
        const RnaV3<T> Rd = Rc.apply(d);
        if (MODE == 0) {
            RnaV3<T> c0, c1, c2; Rc.columns(d, Rd, c0, c1, c2);
            const RnaV3<T> pOc = rnaLd3(vec, c), pAc = rnaLd3(vec + D.offA, c);
            aO.x -= (pOc.x - pOv.x) - (c0.x * pAc.x + c1.x * pAc.y + c2.x * pAc.z);
            aO.y -= (pOc.y - pOv.y) - (c0.y * pAc.x + c1.y * pAc.y + c2.y * pAc.z);
            aO.z -= (pOc.z - pOv.z) - (c0.z * pAc.x + c1.z * pAc.y + c2.z * pAc.z);
        } else {
            const RnaV3<T> Oc = rnaLd3(D.Offset, c);
            aO.x -= (Oc.x - Ov.x) - Rd.x; aO.y -= (Oc.y - Ov.y) - Rd.y; aO.z -= (Oc.z - Ov.z) - Rd.z;
        }
    }
    RnaOut<T> o;
    if (MODE == 2) { o.O = RnaV3<T>{ w2 * aO.x, w2 * aO.x, w2 * aO.x }; o.A = RnaV3<T>{ w2 * aA.x, w2 * aA.y, w2 * aA.z }; }
    else { o.O = RnaV3<T>{ w2 * aO.x, w2 * aO.y, w2 * aO.z }; o.A = RnaV3<T>{ w2 * aA.x, w2 * aA.y, w2 * aA.z }; }
    o.g = T(0);
    if (sv.tgt != T(0)) {   // r_plane and r_conf
        const RnaV3<T> a1O{ D.wf * sv.g * sv.nx, D.wf * sv.g * sv.ny, D.wf * sv.g * sv.nz };
        const T a1g = D.wf * sv.s, a2g = T(-2) * D.wc * sv.g;
        const T m = sv.tgt;
        if (MODE == 2) {
            o.O.x += m * (a1O.x * a1O.x); o.O.y += m * (a1O.y * a1O.y); o.O.z += m * (a1O.z * a1O.z);
            o.g += m * (a1g * a1g + a2g * a2g);
        } else {
            T y1, y2;
            if (MODE == 0) { y1 = a1O.x * pOv.x + a1O.y * pOv.y + a1O.z * pOv.z + a1g * pgv; y2 = a2g * pgv; }
            else { y1 = D.wf * (sv.g * sv.s); y2 = D.wc * (T(1) - sv.g * sv.g); }
            o.O.x += m * (a1O.x * y1); o.O.y += m * (a1O.y * y1); o.O.z += m * (a1O.z * y1);
            o.g += m * (a1g * y1 + a2g * y2);
        }
    }
    (void)w;
    return o;
}

// 1/2 sum of squares of the rows owned by vertex v: its r_plane, r_conf and the r_reg rows of its OUT edges.
// MODEL: rows are F + J delta (computeModelCost + _Graph, solverGPUGaussNewton.t:666-678, :746-756).
template <class T, bool MODEL> RNA_HD T rnaCostAt(const RnaData<T>& D, int v, const T* delta) {
    const RnaPack<T> sv = D.pack[v];
    const RnaRot<T> Rv(sv);
    const RnaV3<T> Ov = rnaLd3(D.Offset, v);
    RnaV3<T> dOv{ T(0), T(0), T(0) }, dAv{ T(0), T(0), T(0) }; T dgv = T(0);
    if (MODEL) { dOv = rnaLd3(delta, v); dAv = rnaLd3(delta + D.offA, v); dgv = delta[D.offG + v]; }
    T e = T(0);
    for (int k = D.outStart[v]; k < D.outStart[v + 1]; ++k) {
        const int b = D.outNbr[k];
        const RnaPack<T> sb_ = D.pack[b];
        const RnaV3<T> Ob = rnaLd3(D.Offset, b);
        const RnaV3<T> d{ sv.ux - sb_.ux, sv.uy - sb_.uy, sv.uz - sb_.uz };
        const RnaV3<T> Rd = Rv.apply(d);
        RnaV3<T> re{ D.w * ((Ov.x - Ob.x) - Rd.x), D.w * ((Ov.y - Ob.y) - Rd.y), D.w * ((Ov.z - Ob.z) - Rd.z) };
        if (MODEL) {
            const RnaV3<T> dOb = rnaLd3(delta, b);
            RnaV3<T> c0, c1, c2; Rv.columns(d, Rd, c0, c1, c2);
            re.x += D.w * ((dOv.x - dOb.x) - (c0.x * dAv.x + c1.x * dAv.y + c2.x * dAv.z));
            re.y += D.w * ((dOv.y - dOb.y) - (c0.y * dAv.x + c1.y * dAv.y + c2.y * dAv.z));
            re.z += D.w * ((dOv.z - dOb.z) - (c0.z * dAv.x + c1.z * dAv.y + c2.z * dAv.z));
        }
        e += rnaDot(re, re);
    }
    if (sv.tgt != T(0)) {
        T r1 = D.wf * (sv.g * sv.s), r2 = D.wc * (T(1) - sv.g * sv.g);
        if (MODEL) {
            r1 += D.wf * sv.g * (sv.nx * dOv.x + sv.ny * dOv.y + sv.nz * dOv.z) + D.wf * sv.s * dgv;
            r2 += T(-2) * D.wc * sv.g * dgv;
        }
        e += sv.tgt * (r1 * r1 + r2 * r2);
    }
    return T(0.5) * e;
}

}  // namespace optb
