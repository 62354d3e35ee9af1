// solver.cu -- generic Gauss-Newton / Levenberg-Marquardt driver with Jacobi-PCG, sm_100a.
//
// Restates the control flow of the reference solver (API/src/solverGPUGaussNewton.t: init :956-1007,
// step :1016-1177) over flat unknown vectors.  Differences that are design, not semantics:
//   * PCG vectors are flat, 64-scalar-aligned concatenations of the unknown arrays; excluded and
//     padding entries hold r = p = Ap = 0, so the streaming kernels (Step2/Step3/update) need no
//     Exclude() test -- they keep zeros at zeros (the reference tests fmap.exclude in every kernel,
//     :371,:424,:450,:539,:554).
//   * reductions are two-level and deterministic (common.cuh) instead of float atomics to one word;
//     the r.z scalar ping-pongs between two slots instead of the per-iteration D2D copy (:1091).
//   * the whole inner PCG loop of a nonlinear iteration is one CUDA graph launch; LM's zeta early-out
//     (:1093-1102, a blocking D2H per PCG step in the reference) is evaluated on the device and turns
//     the remaining captured kernels into no-ops.
//   * PCGStep3 of iteration k is fused into PCGStep1 of iteration k+1 when the energy provides it.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <ctime>
#include <map>
#include "solver.h"
#include "frontend.h"
#include "lm_decide.cuh"

namespace optb {

template <class T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };

template <class T> __device__ __forceinline__ void unpack(const float4& v, T* a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
template <class T> __device__ __forceinline__ void unpack(const double2& v, T* a) { a[0] = v.x; a[1] = v.y; }
__device__ __forceinline__ float4 pack(const float* a) { return make_float4(a[0], a[1], a[2], a[3]); }
__device__ __forceinline__ double2 pack(const double* a) { return make_double2(a[0], a[1]); }

// Multi-GPU: which entries of the flat vectors this rank owns (in VecOf<T>::N-wide units).  Halo entries
// are streamed like everything else (so delta/p stay bitwise consistent with their owner) but must not
// be counted in dot products.  n == 0 means "everything" (single GPU).
struct OwnedRanges {
    long lo[4], hi[4]; int n;
    __device__ __forceinline__ bool owns(long i) const {
        if (n == 0) return true;
        bool o = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) o = o || (k < n && i >= lo[k] && i < hi[k]);
        return o;
    }
};

// Multi-GPU with replicated (global-size) vectors (graph energies): the part of the flat vectors this rank has to
// stream at all = owned entries + halo entries, as up to two contiguous windows (one per unknown array).  n == 0: everything.
struct WindowRanges {
    long lo[3], len[3]; long total; int n;
    __device__ __forceinline__ long map(long j) const {
        if (n == 0 || j < len[0]) return n == 0 ? j : lo[0] + j;
        j -= len[0];
        return (n < 3 || j < len[1]) ? lo[1] + j : lo[2] + (j - len[1]);
    }
};

// Multi-GPU, halo of z without a separate exchange kernel: PCGStep2 stores the z it computes for entries inside this rank's boundary
// segments ALSO into the neighbour's halo entries (NVLink peer stores), and its reduction's last CTA releases the neighbours' flags;
// PCGStep3 forms p on the owned entries first and waits for the incoming flag only before it touches the halo entries.
template <class T> struct ZPush {
    int n;                                   // 0: disabled
    long lo[kMaxHaloSegs], hi[kMaxHaloSegs], shift[kMaxHaloSegs];   // boundary segments in VecOf<T>::N-wide units; destination index = i + shift
    T* dst[kMaxHaloSegs];                    // the neighbour's z vector
    unsigned long long* peerFlag[2];         // neighbours' "halo arrived" words (nullptr: no neighbour on that side)
};
struct HaloWait {
    int n;                                   // 0: disabled
    long lo[8], len[8], total;               // halo entries inside the streamed window, VecOf<T>::N-wide units
    unsigned long long* flag[2];             // my "halo arrived" words, written by the upper / lower neighbour (nullptr: none)
    __device__ __forceinline__ long map(long j) const {
        long i = -1;
#pragma unroll
        for (int k = 0; k < 8; ++k) { if (k < n && i < 0) { if (j < len[k]) i = lo[k] + j; else j -= len[k]; } }
        return i;
    }
};

// ------------------------------------------------------------------------------------ PCGStep2
// solverGPUGaussNewton.t:446-489.  alpha = rz_old / pAp (guarded), delta += alpha p, r -= alpha Ap,
// z = pre*r, rz_new = z.r ; LM: q = 1/2 delta.(r+b).
template <class T, bool LM, bool USEPRE>
__global__ void __launch_bounds__(kBlock) k_pcgStep2(long nVec, WindowRanges win, ReduceWorkspace ws, OwnedRanges own, int rzOld, int rzNew,
                                                     T* __restrict__ delta, const T* __restrict__ p, T* __restrict__ r,
                                                     const T* __restrict__ Ap, const T* __restrict__ pre,
                                                     const T* __restrict__ b, T* __restrict__ z, ZPush<T> push, int zetaIter, T zetaTol) {
    using V = typename VecOf<T>::type; constexpr int N = VecOf<T>::N;
    pdlTrigger(); pdlWait();
    if (LM && ws.scalars[kFlags] != 0.0) return;
    // The thread's first batch of loads is issued BEFORE alpha is awaited: none of the operands depends on it (they were all written
    // by earlier kernels on this stream), and on several GPUs alpha = rz / p.Ap arrives through the NVLink mailboxes a few
    // microseconds after the kernel starts -- that wait now overlaps the first HBM round trip instead of preceding it.
    const long stride = (long)gridDim.x * kBlock;
    long j = (long)blockIdx.x * kBlock + threadIdx.x;
    V vd{}, vp{}, vr{}, vap{}, vpr{}, vb{};
    long i = 0;
    if (j < nVec) {
        i = win.map(j);
        vd = reinterpret_cast<const V*>(delta)[i]; vp = reinterpret_cast<const V*>(p)[i]; vr = reinterpret_cast<const V*>(r)[i];
        vap = reinterpret_cast<const V*>(Ap)[i];
        if (USEPRE) vpr = reinterpret_cast<const V*>(pre)[i];
        if (LM) vb = reinterpret_cast<const V*>(b)[i];
    }
    double dnum, dden; loadScalars2(ws, rzOld, kPAp, dnum, dden);
    const T num = (T)dnum, den = (T)dden;
    T alpha = T(0);
    if (den > T(0)) alpha = num / den;
    T accRz = T(0), accQ = T(0);
    while (j < nVec) {
        const bool mine = own.owns(i);
        T d[N], pp[N], rr[N], ap[N], pr[N], bb[N], zz[N];
        unpack(vd, d); unpack(vp, pp); unpack(vr, rr); unpack(vap, ap);
        if (USEPRE) unpack(vpr, pr);
        if (LM) unpack(vb, bb);
        const long iCur = i;
        j += stride;
        if (j < nVec) {   // next batch in flight while this one is computed and stored
            i = win.map(j);
            vd = reinterpret_cast<const V*>(delta)[i]; vp = reinterpret_cast<const V*>(p)[i]; vr = reinterpret_cast<const V*>(r)[i];
            vap = reinterpret_cast<const V*>(Ap)[i];
            if (USEPRE) vpr = reinterpret_cast<const V*>(pre)[i];
            if (LM) vb = reinterpret_cast<const V*>(b)[i];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            d[k] = d[k] + alpha * pp[k];
            rr[k] = rr[k] - alpha * ap[k];
            zz[k] = USEPRE ? pr[k] * rr[k] : rr[k];
            if (mine) {
                accRz += zz[k] * rr[k];
                if (LM) accQ += T(0.5) * (d[k] * (rr[k] + bb[k]));
            }
        }
        reinterpret_cast<V*>(delta)[iCur] = pack(d);
        reinterpret_cast<V*>(r)[iCur] = pack(rr);
        if (mine) {   // halo entries of z belong to the neighbour's NVLink store
            reinterpret_cast<V*>(z)[iCur] = pack(zz);
            if (push.n) {
#pragma unroll
                for (int k = 0; k < kMaxHaloSegs; ++k) if (k < push.n && iCur >= push.lo[k] && iCur < push.hi[k]) reinterpret_cast<V*>(push.dst[k])[iCur + push.shift[k]] = pack(zz);
            }
        }
    }
    if (!LM) { gridReduceTo<kBlock>((double)accRz, ws, rzNew, push.n ? push.peerFlag : nullptr); return; }
    // LM: r.z and q in one arrival round; on a single GPU the CTA that finishes the reduction also takes the zeta early-out decision
    // (:1093-1102) -- the reduced q is in its hands -- instead of a one-thread kernel per PCG step.  The flag records the breaking
    // iteration (1 + l): PCGStep3 of THAT iteration still runs (the reference updates p before it tests zeta), everything later is a no-op.
    __shared__ double s_q;
    double q1 = 0.0;
    const bool last = gridReduceTo2<kBlock>((double)accRz, (double)accQ, ws, rzNew, kQ, push.n ? push.peerFlag : nullptr, nullptr, &s_q);
    if (last && zetaIter >= 0 && threadIdx.x == 0) {
        q1 = s_q;
        const T Q1 = (T)q1, Q0 = (T)ws.scalars[kQPrev];
        const T zeta = T(zetaIter + 1) * (Q1 - Q0) / Q1;
        if (zeta < zetaTol) ws.scalars[kFlags] = 1.0 + zetaIter;
        ws.scalars[kQPrev] = (double)Q1;
    }
}

// PCGStep2_1stHalf (:491-503): delta += alpha p
template <class T>
__global__ void __launch_bounds__(kBlock) k_pcgStep2FirstHalf(long nVec, WindowRanges win, ReduceWorkspace ws, int rzOld, T* __restrict__ delta, const T* __restrict__ p) {
    using V = typename VecOf<T>::type; constexpr int N = VecOf<T>::N;
    if (ws.scalars[kFlags] != 0.0) return;
    double dnum, dden; loadScalars2(ws, rzOld, kPAp, dnum, dden);
    const T num = (T)dnum, den = (T)dden;
    T alpha = T(0);
    if (den > T(0)) alpha = num / den;
    for (long j = (long)blockIdx.x * kBlock + threadIdx.x; j < nVec; j += (long)gridDim.x * kBlock) {
        const long i = win.map(j);
        T d[N], pp[N];
        unpack(reinterpret_cast<const V*>(delta)[i], d); unpack(reinterpret_cast<const V*>(p)[i], pp);
#pragma unroll
        for (int k = 0; k < N; ++k) d[k] = d[k] + alpha * pp[k];
        reinterpret_cast<V*>(delta)[i] = pack(d);
    }
}
// PCGStep2_2ndHalf (:505-534): r = b - Adelta, z = pre*r, rz_new, q
template <class T, bool USEPRE>
__global__ void __launch_bounds__(kBlock) k_pcgStep2SecondHalf(long nVec, WindowRanges win, ReduceWorkspace ws, OwnedRanges own, int rzNew, const T* __restrict__ delta,
                                                               T* __restrict__ r, const T* __restrict__ Adelta, const T* __restrict__ pre,
                                                               const T* __restrict__ b, T* __restrict__ z) {
    using V = typename VecOf<T>::type; constexpr int N = VecOf<T>::N;
    if (ws.scalars[kFlags] != 0.0) return;
    T accRz = T(0), accQ = T(0);
    for (long j = (long)blockIdx.x * kBlock + threadIdx.x; j < nVec; j += (long)gridDim.x * kBlock) {
        const long i = win.map(j);
        T d[N], ad[N], pr[N], bb[N], rr[N], zz[N];
        unpack(reinterpret_cast<const V*>(delta)[i], d); unpack(reinterpret_cast<const V*>(Adelta)[i], ad);
        unpack(reinterpret_cast<const V*>(b)[i], bb);
        if (USEPRE) unpack(reinterpret_cast<const V*>(pre)[i], pr);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            rr[k] = bb[k] - ad[k];
            zz[k] = USEPRE ? pr[k] * rr[k] : rr[k];
            if (own.owns(i)) {
                accRz += zz[k] * rr[k];
                accQ += T(0.5) * (d[k] * (rr[k] + bb[k]));
            }
        }
        reinterpret_cast<V*>(r)[i] = pack(rr); if (own.owns(i)) reinterpret_cast<V*>(z)[i] = pack(zz);
    }
    gridReduceTo<kBlock>((double)accRz, ws, rzNew);
    gridReduceTo<kBlock>((double)accQ, ws, kQ);
}

// ------------------------------------------------------------------------------------ PCGStep3
// :537-550.  beta = rz_new / rz_old (guarded), p = z + beta p
template <class T, bool LM>
__global__ void __launch_bounds__(kBlock) k_pcgStep3(long nVec, WindowRanges win, ReduceWorkspace ws, int rzNew, int rzOld, const T* __restrict__ z, T* __restrict__ p,
                                                     OwnedRanges own, HaloWait hw, int lIter) {
    using V = typename VecOf<T>::type; constexpr int N = VecOf<T>::N;
    pdlTrigger(); pdlWait();
    if (LM) { const double f = ws.scalars[kFlags]; if (f != 0.0 && f != 1.0 + lIter) return; }   // a flag raised by THIS iteration's zeta test still lets its Step3 run
    double drn, dro; loadScalars2(ws, rzNew, rzOld, drn, dro);
    const T rn = (T)drn, ro = (T)dro;
    T beta = T(0);
    if (ro > T(0)) beta = rn / ro;
    auto one = [&](long i) {
        T zz[N], pp[N];
        unpack(reinterpret_cast<const V*>(z)[i], zz); unpack(reinterpret_cast<const V*>(p)[i], pp);
#pragma unroll
        for (int k = 0; k < N; ++k) pp[k] = zz[k] + beta * pp[k];
        reinterpret_cast<V*>(p)[i] = pack(pp);
    };
    for (long j = (long)blockIdx.x * kBlock + threadIdx.x; j < nVec; j += (long)gridDim.x * kBlock) {
        const long i = win.map(j);
        if (hw.n && !own.owns(i)) continue;      // halo entries: second pass, once the neighbours' z has arrived
        one(i);
    }
    if (hw.n) {
        if (threadIdx.x == 0) {
            const unsigned long long e = haloEpochNow(ws);
            for (int d = 0; d < 2; ++d) if (hw.flag[d]) { SpinGuard g; while (ldAcquireSys(hw.flag[d]) < e) g.tick(); }
        }
        __syncthreads();
        for (long j = (long)blockIdx.x * kBlock + threadIdx.x; j < hw.total; j += (long)gridDim.x * kBlock) {
            const long i = hw.map(j);
            T zz[N], pp[N];   // z was written by another GPU: read it past L1
            unpack(__ldcg(reinterpret_cast<const V*>(z) + i), zz); unpack(reinterpret_cast<const V*>(p)[i], pp);
#pragma unroll
            for (int k = 0; k < N; ++k) pp[k] = zz[k] + beta * pp[k];
            reinterpret_cast<V*>(p)[i] = pack(pp);
        }
    }
}

// LM zeta early-out on the device (:1093-1102): zeta = (l+1)(Q1-Q0)/Q1 < q_tolerance => stop iterating
template <class T>
__global__ void k_lmZeta(ReduceWorkspace ws, int lIter, T qTol) {
    if (ws.scalars[kFlags] != 0.0) return;
    T Q1 = (T)loadScalarGlobal(ws, kQ), Q0 = (T)ws.scalars[kQPrev];
    T zeta = T(lIter + 1) * (Q1 - Q0) / Q1;
    if (zeta < qTol) ws.scalars[kFlags] = 1.0 + lIter;   // remembers the breaking iteration
    ws.scalars[kQPrev] = (double)Q1;
}

// ------------------------------------------------------------------------------------ small vector kernels
template <class T> __global__ void __launch_bounds__(kBlock) k_axpyInto(long n, T* __restrict__ x, const T* __restrict__ d) {   // PCGLinearUpdate :552-557
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long)gridDim.x * kBlock) x[i] = x[i] + d[i];
}
// LM: savePreviousUnknowns + PCGLinearUpdate in one pass (:1108-1113, :552-557)
template <class T> __global__ void __launch_bounds__(kBlock) k_saveAndUpdate(long n, T* __restrict__ saved, T* __restrict__ x, const T* __restrict__ d) {
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long)gridDim.x * kBlock) { const T v = x[i]; saved[i] = v; x[i] = v + d[i]; }
}
// ---- LM accept / reject / trust-region update on the device (OPTB_DEVICE_LM=1; Plan::runDeviceIterations).  One thread: the decision
// is a dozen scalar operations (lm_decide.cuh = :1119-1157); what it buys is that no host round trip separates two nonlinear iterations.
template <class T>
__global__ void k_lmDecide(ReduceWorkspace ws, LmLimits<T> lim) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double* s = ws.scalars;
    if (s[kLmGate] != 0.0) { s[kLmReject] = 0.0; return; }   // the solve ended in an earlier enqueued iteration
    LmState<T> st{ (T)s[kLmPrevCost], (T)s[kRadius], (T)s[kLmFactor] };
    const int v = lmDecide<T>((T)s[kCost], (T)s[kModelCost], st, lim);
    s[kLmPrevCost] = (double)st.prevCost; s[kRadius] = (double)st.radius; s[kLmFactor] = (double)st.decreaseFactor;
    s[kLmReject] = (v == kLmRejected || v == kLmRadiusTooSmall) ? 1.0 : 0.0;
    if (v == kLmFunctionTolerance || v == kLmRadiusTooSmall) s[kLmGate] = (double)v;   // the host's step() returns 0 here, before ++nIter
    else s[kLmIters] += 1.0;
}
template <class T> __global__ void __launch_bounds__(kBlock) k_copyIfRejected(long n, T* __restrict__ dst, const T* __restrict__ src, ReduceWorkspace ws) {   // revertUpdate :559-564
    if (ws.scalars[kLmReject] == 0.0) return;
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long)gridDim.x * kBlock) dst[i] = src[i];
}
template <class T> __global__ void __launch_bounds__(kBlock) k_copy(long n, T* __restrict__ dst, const T* __restrict__ src) {
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long)gridDim.x * kBlock) dst[i] = src[i];
}
// PCGFinalizeDiagonal (:631-664).  `pre` holds guardedInvert(diag) from PCGInit1 on entry (0 on
// excluded/padding entries, which stay inert), the true LM preconditioner on exit.
template <class T>
__global__ void __launch_bounds__(kBlock) k_lmFinalizeDiagonal(long n, WindowRanges win, ReduceWorkspace ws, OwnedRanges own, int rzSlot, T* __restrict__ CtC, T* __restrict__ SSq,
                                                               T* __restrict__ pre, const T* __restrict__ r, T* __restrict__ b, T* __restrict__ p,
                                                               const T* __restrict__ delta, T minDiag, T maxDiag) {
    // the trust-region radius of this nonlinear iteration lives in the plan's scalar block (written by the host before the iteration): the
    // captured iteration graph is the same for every radius.  CtC arrives unscaled (diag J^T J) and gets its 1/radius here (:616-622).
    const T radius = (T)ws.scalars[kRadius], invRadius = T(1) / radius;
    T accD = T(0), accQ = T(0);
    for (long j = (long)blockIdx.x * kBlock + threadIdx.x; j < n; j += (long)gridDim.x * kBlock) {
        const long i = win.map(j);
        // active NOW = PCGInit1 left a non-zero guardedInvert(diag) in pre (0 on excluded / padding / halo-of-nothing entries).  SSq is saved
        // at the first nonlinear iteration only (:1038-1041); an unknown whose Exclude() mask cleared later gets its scale now.
        const T preIn = pre[i];
        T ssq = SSq[i];
        if (preIn > T(0) && !(ssq > T(0))) { ssq = preIn; SSq[i] = ssq; }
        if (preIn > T(0) && ssq > T(0)) {
            T unclamped = CtC[i] * invRadius;
            T invS = T(1) / ssq;
            T clampMul = invS / radius;
            T mn = minDiag * clampMul, mx = maxDiag * clampMul;
            T c = fmin(fmax(unclamped, mn), mx);
            CtC[i] = c;
            T pr = T(1) / (c + radius * unclamped);
            pre[i] = pr;
            T res = r[i];
            b[i] = res;
            T pp = pr * res;
            if (own.owns(i)) p[i] = pp;   // halo entries of p arrive from the neighbour
            if (own.owns(i)) {
                accD += res * pp;
                accQ += T(0.5) * (delta[i] * (res + res));
            }
        } else {
            CtC[i] = T(0); pre[i] = T(0); b[i] = T(0); if (own.owns(i)) p[i] = T(0);
        }
    }
    gridReduceTo<kBlock>((double)accQ, ws, kQ);
    gridReduceTo<kBlock>((double)accD, ws, rzSlot);
}
__global__ void k_setScalar(ReduceWorkspace ws, int slot, double v) { ws.scalars[slot] = v; }
__global__ void k_copyScalar(ReduceWorkspace ws, int dst, int src) { ws.scalars[dst] = loadScalarGlobal(ws, src); }
// multi-GPU: materialise the globally reduced values into the local scalar block before a host read-back
__global__ void k_finalizeScalars(ReduceWorkspace ws, unsigned mask) {
    const int s = threadIdx.x;
    if (s < kNumScalars && (mask >> s) & 1u) ws.scalars[s] = loadScalarGlobal(ws, s);
}
// Halo exchange of one flat vector: a single CTA stores this rank's boundary segments
// into the neighbours' halo segments over NVLink (peer pointers from CUDA IPC), releases a flag at each
// neighbour and then waits for the neighbours' flags, so the next kernel on the stream sees fresh halos.
template <class T>
__global__ void __launch_bounds__(1024) k_haloExchange(const HaloPlan* __restrict__ hpp, ReduceWorkspace ws, const T* __restrict__ vec) {
    const HaloPlan& hp = *hpp;
    for (int s = 0; s < hp.nSeg; ++s) {
        T* dst = reinterpret_cast<T*>(hp.peerVec[hp.dir[s]]);
        if (!dst) continue;
        for (long i = threadIdx.x; i < hp.len[s]; i += blockDim.x) dst[hp.dstOff[s] + i] = vec[hp.srcOff[s] + i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        // own sequence number and own flag words: several exchanges may happen between two r.z productions (p before the loop kernel,
        // delta after it), and the in-kernel halo paths number THEIR flags by the r.z epochs
        const unsigned long long e = *hp.epoch + 1ull; *hp.epoch = e;
        for (int d = 0; d < 2; ++d) if (hp.peerVec[d]) stReleaseSys(hp.peerXFlag[d], e);
        for (int d = 0; d < 2; ++d) if (hp.peerVec[d]) { SpinGuard g; while (ldAcquireSys(hp.myXFlag[d]) < e) g.tick(); }
    }
}

// ====================================================================================== Plan<T>
template <class T>
struct Plan : PlanBase {
    EnergyOps<T>* E;
    bool lm;
    SolverParameters sp;
    int verbosity, collectTiming;
    long n = 0;                 // padded unknown count
    T *delta = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *Ap = nullptr, *pre = nullptr;
    T *p2 = nullptr, *pFinal = nullptr;   // second direction buffer for the fused Step3+Step1 path; where p lives after the loop
    T *b = nullptr, *Adelta = nullptr, *CtC = nullptr, *SSq = nullptr, *prevX = nullptr;   // LM only
    ReduceWorkspace ws{};
    OwnedRanges ownVec{}, ownScalar{};       // n == 0: single GPU, everything owned
    WindowRanges winVec{}, winScalar{};      // n == 0: stream the whole vectors
    long nVecWin() const { return winVec.n ? winVec.total : n / VecOf<T>::N; }
    // multi-GPU mailbox block (cudaMalloc'ed so it can be exported through CUDA IPC), see common.cuh CommView
    struct CommBlock { unsigned long long words[kNumScalars][2][kMaxRanks][2]; unsigned long long haloFlags[2]; unsigned long long epoch[kNumScalars]; unsigned long long haloEpoch;
                       unsigned long long fitFlags[kMaxRanks]; unsigned long long xFlags[2]; };   // fitFlags[r]: (connect generation << 2) | 1 (rank r cannot run the collective loop kernel) or 2 (it can)
    CommBlock* commBlock = nullptr;
    int world = 1, rank = 0;
    HaloPlan haloP{}, haloZ{}, haloD{};   // halo descriptions of p, z and delta (same segments, different peer vectors)
    ZPush<T> zPush{}; HaloWait zWait{};   // in-kernel halo of z (Step2 pushes, Step3 waits); n == 0: not available, use k_haloExchange
    HaloPlan* dHaloP = nullptr; HaloPlan* dHaloZ = nullptr; HaloPlan* dHaloD = nullptr;   // device copies (kernels take them by pointer)
    unsigned long long** dPeerTable = nullptr;
    std::vector<void*> ipcOpened;
    double* hScalars = nullptr;  // pinned mirror
    cudaStream_t stream = nullptr, userStream = nullptr;
    bool ownStream = true;
    cudaEvent_t evEntry = nullptr, evLoop0 = nullptr, evLoop1 = nullptr, evCost = nullptr, evExit = nullptr;
    mutable bool costPending = false, loopMsPending = false;   // GN steps do not wait for their own cost: resolved on demand
    int numSMs = 148;
    int gridStream = 0;          // CTAs for the streaming kernels
    int gridStep2 = 0;           // CTAs for PCGStep2 (co-resident count)
    cudaGraphExec_t loopExec = nullptr; unsigned long long loopExecKey = 0; int graphBuilds = 0;
    bool asyncSteps = true;      // GN: Opt_ProblemStep returns without waiting for its cost (OPTB_SYNC_STEPS=1: block like the reference)
    bool useGraph = true, useFusion = true, usePdl = false;   // PDL measured slower here (early-resident dependents crowd the streaming kernels): opt-in via OPTB_PDL=1
    // nonlinear state (:994-1006)
    int nIter = 0;
    T prevCost = 0;
    T trust_region_radius = 0, radius_decrease_factor = 0, min_lm_diagonal = 0, max_lm_diagonal = 0;
    int maxLinIters = -1;
    int launches = 0; float loopMs = 0;
    bool finished = false, reported = false;

    Plan(EnergyOps<T>* e, bool isLM, int verb, int timing) : E(e), lm(isLM), verbosity(verb), collectTiming(timing) {
        int dev = 0; OPT_CUDA_CHECK(cudaGetDevice(&dev));
        OPT_CUDA_CHECK(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev));
        n = E->layoutUnknowns();
        auto alloc = [&](T** v) { OPT_CUDA_CHECK(cudaMalloc((void**)v, n * sizeof(T))); OPT_CUDA_CHECK(cudaMemset(*v, 0, n * sizeof(T))); };
        alloc(&delta); alloc(&r); alloc(&z); alloc(&p); alloc(&Ap); alloc(&pre); alloc(&p2); pFinal = p;
        if (lm) { alloc(&b); alloc(&Adelta); alloc(&CtC); alloc(&SSq); alloc(&prevX); }
        OPT_CUDA_CHECK(cudaMalloc((void**)&ws.partials, sizeof(double) * kNumScalars * kMaxPartials));
        OPT_CUDA_CHECK(cudaMalloc((void**)&ws.counters, sizeof(unsigned) * kNumScalars));
        OPT_CUDA_CHECK(cudaMalloc((void**)&ws.scalars, sizeof(double) * kNumScalars));
        OPT_CUDA_CHECK(cudaMemset(ws.counters, 0, sizeof(unsigned) * kNumScalars));
        OPT_CUDA_CHECK(cudaMemset(ws.scalars, 0, sizeof(double) * kNumScalars));
        OPT_CUDA_CHECK(cudaMalloc((void**)&commBlock, sizeof(CommBlock) + kPeerScratchBytes));   // mailboxes, then the peer scratch area
        OPT_CUDA_CHECK(cudaMemset(commBlock, 0, sizeof(CommBlock) + kPeerScratchBytes));
        ws.comm.world = 1; ws.comm.rank = 0;
        OPT_CUDA_CHECK(cudaMallocHost((void**)&hScalars, sizeof(double) * kNumScalars));
        OPT_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        OPT_CUDA_CHECK(cudaEventCreateWithFlags(&evEntry, cudaEventDisableTiming));
        OPT_CUDA_CHECK(cudaEventCreateWithFlags(&evCost, cudaEventDisableTiming)); OPT_CUDA_CHECK(cudaEventCreateWithFlags(&evExit, cudaEventDisableTiming));
        OPT_CUDA_CHECK(cudaEventCreate(&evLoop0)); OPT_CUDA_CHECK(cudaEventCreate(&evLoop1));
        long nVec = n / VecOf<T>::N;
        gridStream = (int)std::min<long>((nVec + kBlock - 1) / kBlock, (long)numSMs * 8);
        if (gridStream < 1) gridStream = 1;
        {   // PCGStep2 keeps more registers than the 32 that 8 CTAs/SM would allow: size its grid to what is actually co-resident, so the
            // grid-stride loop runs as ONE wave (a partial second wave of 2 CTAs per SM cannot saturate HBM)
            const void* fn = lm ? (E->usePreconditioner ? (const void*)k_pcgStep2<T, true, true> : (const void*)k_pcgStep2<T, true, false>)
                                : (E->usePreconditioner ? (const void*)k_pcgStep2<T, false, true> : (const void*)k_pcgStep2<T, false, false>);
            int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBlock, 0);
            if (occ < 1) occ = 1; if (occ > 8) occ = 8;
            gridStep2 = (int)std::min<long>((nVec + kBlock - 1) / kBlock, (long)numSMs * occ);
            if (gridStep2 < 1) gridStep2 = 1;
            if (const char* e = getenv("OPTB_GRID_STEP2")) gridStep2 = std::max(1, std::min(atoi(e), kMaxPartials));   // experiments: also changes the summation order of r.z
        }
        if (const char* s = getenv("OPTB_NO_GRAPH")) useGraph = atoi(s) == 0;
        if (const char* s = getenv("OPTB_NO_FUSION")) useFusion = atoi(s) == 0;
        if (const char* s = getenv("OPTB_PDL")) usePdl = atoi(s) != 0;
        persistentMode = -1; if (const char* s = getenv("OPTB_PERSISTENT")) persistentMode = atoi(s) != 0 ? 1 : 0;
        if (const char* s = getenv("OPTB_SYNC_STEPS")) asyncSteps = atoi(s) == 0;
        if (getenv("OPTB_NO_ITER_GRAPH")) useIterGraph = false;
        // The reference's harness always asks for the timing table (examples/shared/OptSolver.h:47-48), so the flag alone must
        // not cost the fast path: by default the table has one row per REGION (PCGInit1, the whole PCG loop, update, cost);
        // OPTB_TIMING_TABLE=1 switches to one row per kernel by issuing the per-step kernels individually.
        if (collectTiming) { timingOn = true; if (getenv("OPTB_TIMING_TABLE")) { timingFine = true; useGraph = false; useFusion = false; persistentMode = 0; } }
        useOnChipLoop = persistentMode != 0;   // OPTB_PERSISTENT=0: kernel-per-step loops only
        if (verbosity > 1) printf("opt_b200: plan %s %s %s: %ld unknowns (padded), %d SMs\n", E->name(), lm ? "LM" : "GN", sizeof(T) == 8 ? "fp64" : "fp32", n, numSMs);
    }
    ~Plan() override {
        cudaStreamSynchronize(stream);
        dropTimedLaunches();
        if (loopExec) cudaGraphExecDestroy(loopExec);
        if (iterExec) cudaGraphExecDestroy(iterExec);
        if (devIterExec) cudaGraphExecDestroy(devIterExec);
        if (hLmState) cudaFreeHost(hLmState);
        if (hRadius) cudaFreeHost(hRadius);
        for (T* v : { delta, r, z, p, p2, Ap, pre, b, Adelta, CtC, SSq, prevX }) if (v) cudaFree(v);
        for (void* q : ipcOpened) cudaIpcCloseMemHandle(q);
        cudaFree(commBlock); cudaFree(dHaloP); cudaFree(dHaloZ); cudaFree(dHaloD); cudaFree(dPeerTable);
        cudaFree(ws.partials); cudaFree(ws.counters); cudaFree(ws.scalars); cudaFreeHost(hScalars);
        cudaEventDestroy(evEntry); cudaEventDestroy(evLoop0); cudaEventDestroy(evLoop1); cudaEventDestroy(evCost); cudaEventDestroy(evExit);
        if (ownStream) cudaStreamDestroy(stream);
        delete E;
    }
    int ctxGate = 0;   // LaunchCtx::gate of the launches being enqueued (non-zero only while runDeviceIterations captures its graph)
    LaunchCtx ctx() const { return LaunchCtx{ stream, ws, numSMs, usePdl, ctxGate }; }

    void setParameter(const char* name, const void* v) override {   // :1205-1221
        std::string s(name);
#define SPF(x) if (s == #x) { sp.x = *(const float*)v; return; }
#define SPI(x) if (s == #x) { sp.x = *(const int*)v; return; }
        SPF(min_relative_decrease) SPF(min_trust_region_radius) SPF(max_trust_region_radius) SPF(q_tolerance)
        SPF(function_tolerance) SPF(trust_region_radius) SPF(radius_decrease_factor) SPF(min_lm_diagonal) SPF(max_lm_diagonal)
        SPI(residual_reset_period) SPI(nIterations) SPI(lIterations)
#undef SPF
#undef SPI
        if (verbosity > 0) printf("Warning: tried to set nonexistent solver parameter %s\n", name);
    }
    void setStream(cudaStream_t s) override { userStream = s; }
    void setMaxLinearIterations(int k) override { maxLinIters = k; }
    int launchesLastStep() const override { return launches; }
    float pcgLoopMsLastStep() const override {
        if (loopMsPending) { OPT_CUDA_CHECK(cudaEventSynchronize(evLoop1)); float ms = 0; OPT_CUDA_CHECK(cudaEventElapsedTime(&ms, evLoop0, evLoop1)); const_cast<Plan*>(this)->loopMs = ms; loopMsPending = false; }
        return loopMs;
    }
    long numUnknowns() const override { return n; }
    int unknownLayout(int maxArrays, long* offsets, long* elements, int* channels) const override {
        int k = 0;
        for (auto& u : E->unknowns) { if (k < maxArrays) { offsets[k] = u.offset; elements[k] = u.elements; channels[k] = u.channels; } ++k; }
        return k;
    }
    double currentCost() const override { const_cast<Plan*>(this)->resolveCost(); return (double)prevCost; }
    // Gauss-Newton takes no decision from the cost, so Opt_ProblemStep does not wait for it (the reference blocks on a D2H copy per
    // nonlinear iteration, solverGPUGaussNewton.t:790-797): the value lands in pinned memory behind an event and is picked up by
    // Opt_ProblemCurrentCost, by the next verbose print, or when the solve ends.
    void resolveCost() {
        if (!costPending) return;
        OPT_CUDA_CHECK(cudaEventSynchronize(evCost));
        prevCost = (T)hScalars[kCost]; costPending = false;
    }
    void computeCostAsync() {
        E->cost(ctx()); ++launches;
        if (world > 1) { k_finalizeScalars<<<1, 32, 0, stream>>>(ws, (1u << kRz0) | (1u << kRz1) | (1u << kPAp) | (1u << kCost) | (1u << kModelCost) | (1u << kQ)); ++launches; }
        OPT_CUDA_CHECK(cudaMemcpyAsync(hScalars, ws.scalars, sizeof(double) * kNumScalars, cudaMemcpyDeviceToHost, stream));
        OPT_CUDA_CHECK(cudaEventRecord(evCost, stream));
        costPending = true;
        // whatever the caller queues next on ITS stream (default: the NULL stream) comes after this step's work
        OPT_CUDA_CHECK(cudaEventRecord(evExit, stream));
        OPT_CUDA_CHECK(cudaStreamWaitEvent(userStream, evExit, 0));
    }

    // order our stream after whatever the caller queued on its stream (default: the legacy NULL stream,
    // which is what the reference runs on, util.t:828)
    void enter() {
        OPT_CUDA_CHECK(cudaEventRecord(evEntry, userStream));
        OPT_CUDA_CHECK(cudaStreamWaitEvent(stream, evEntry, 0));
    }
    // ---- multi-GPU wiring (include/OptB200.h: OptB200_PlanCommHandles / OptB200_PlanConnect)
    static constexpr int kHandlesPerRank = 4;   // mailbox block (+ peer scratch), z, p, delta
    int commHandles(void* out, int outBytes) override {
        if (outBytes < kHandlesPerRank * (int)sizeof(cudaIpcMemHandle_t)) return -1;
        cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)out;
        OPT_CUDA_CHECK(cudaIpcGetMemHandle(&h[0], commBlock));
        OPT_CUDA_CHECK(cudaIpcGetMemHandle(&h[1], z));
        OPT_CUDA_CHECK(cudaIpcGetMemHandle(&h[2], p));
        OPT_CUDA_CHECK(cudaIpcGetMemHandle(&h[3], delta));
        return kHandlesPerRank * (int)sizeof(cudaIpcMemHandle_t);
    }
    int connect(int rank_, int world_, const void* allHandles, int ownedBegin, int ownedEnd, int upperRows, int lowerRows) override {
        if (world_ < 1 || world_ > kMaxRanks || rank_ < 0 || rank_ >= world_) return -1;
        if (!E->setSlab(ownedBegin, ownedEnd, upperRows, lowerRows)) { fprintf(stderr, "opt_b200: energy %s has no slab decomposition\n", E->name()); return -2; }
        OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
        world = world_; rank = rank_;
        const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)allHandles;
        auto open = [&](const cudaIpcMemHandle_t& hh) { void* q = nullptr; OPT_CUDA_CHECK(cudaIpcOpenMemHandle(&q, hh, cudaIpcMemLazyEnablePeerAccess)); ipcOpened.push_back(q); return q; };
        ws.comm.world = world; ws.comm.rank = rank; ws.comm.epoch = commBlock->epoch;
        for (int r = 0; r < world; ++r) {
            CommBlock* cb = (r == rank) ? commBlock : (CommBlock*)open(h[kHandlesPerRank * r + 0]);
            hostPeerTable[r] = &cb->words[0][0][0][0];
            peerBlocks[r] = cb;
        }
        if (!dPeerTable) OPT_CUDA_CHECK(cudaMalloc((void**)&dPeerTable, sizeof(void*) * kMaxRanks));
        OPT_CUDA_CHECK(cudaMemcpy(dPeerTable, hostPeerTable, sizeof(void*) * kMaxRanks, cudaMemcpyHostToDevice));
        ws.comm.peerWords = dPeerTable;
        void* upZ = nullptr; void* upP = nullptr; void* dnZ = nullptr; void* dnP = nullptr; void* upD = nullptr; void* dnD = nullptr;
        if (rank > 0 && upperRows > 0) { upZ = open(h[kHandlesPerRank * (rank - 1) + 1]); upP = open(h[kHandlesPerRank * (rank - 1) + 2]); upD = open(h[kHandlesPerRank * (rank - 1) + 3]); }
        if (rank + 1 < world && lowerRows > 0) { dnZ = open(h[kHandlesPerRank * (rank + 1) + 1]); dnP = open(h[kHandlesPerRank * (rank + 1) + 2]); dnD = open(h[kHandlesPerRank * (rank + 1) + 3]); }
        auto segs = E->haloSegments();
        auto fill = [&](HaloPlan& hp, void* up, void* dn) {
            hp = HaloPlan{};
            hp.nSeg = (int)std::min<size_t>(segs.size(), (size_t)kMaxHaloSegs);
            for (int i = 0; i < hp.nSeg && i < kMaxHaloSegs; ++i) { hp.srcOff[i] = segs[i].srcOff; hp.dstOff[i] = segs[i].dstOff; hp.len[i] = segs[i].len; hp.dir[i] = segs[i].dir; }
            hp.peerVec[0] = up; hp.peerVec[1] = dn;
            // my data lands in the upper neighbour as "from below" (its haloFlags[1]) and in the lower neighbour as "from above" (its haloFlags[0])
            hp.peerFlag[0] = up ? &peerBlocks[rank - 1]->haloFlags[1] : nullptr;
            hp.peerFlag[1] = dn ? &peerBlocks[rank + 1]->haloFlags[0] : nullptr;
            hp.myFlag[0] = &commBlock->haloFlags[0]; hp.myFlag[1] = &commBlock->haloFlags[1];
            hp.peerXFlag[0] = up ? &peerBlocks[rank - 1]->xFlags[1] : nullptr;
            hp.peerXFlag[1] = dn ? &peerBlocks[rank + 1]->xFlags[0] : nullptr;
            hp.myXFlag[0] = &commBlock->xFlags[0]; hp.myXFlag[1] = &commBlock->xFlags[1];
            hp.epoch = &commBlock->haloEpoch;
        };
        fill(haloP, upP, dnP); fill(haloZ, upZ, dnZ); fill(haloD, upD, dnD);
        if (!dHaloP) { OPT_CUDA_CHECK(cudaMalloc((void**)&dHaloP, sizeof(HaloPlan))); OPT_CUDA_CHECK(cudaMalloc((void**)&dHaloZ, sizeof(HaloPlan))); OPT_CUDA_CHECK(cudaMalloc((void**)&dHaloD, sizeof(HaloPlan))); }
        OPT_CUDA_CHECK(cudaMemcpy(dHaloP, &haloP, sizeof(HaloPlan), cudaMemcpyHostToDevice));
        OPT_CUDA_CHECK(cudaMemcpy(dHaloZ, &haloZ, sizeof(HaloPlan), cudaMemcpyHostToDevice));
        OPT_CUDA_CHECK(cudaMemcpy(dHaloD, &haloD, sizeof(HaloPlan), cudaMemcpyHostToDevice));
        auto ranges = E->ownedScalarRanges();
        ownVec = OwnedRanges{}; ownScalar = OwnedRanges{};
        for (size_t i = 0; i < ranges.size() && i < 4; ++i) {
            ownScalar.lo[i] = ranges[i].first; ownScalar.hi[i] = ranges[i].second;
            if (ranges[i].first % VecOf<T>::N || ranges[i].second % VecOf<T>::N) { fprintf(stderr, "opt_b200: slab boundary not vector aligned\n"); return -3; }
            ownVec.lo[i] = ranges[i].first / VecOf<T>::N; ownVec.hi[i] = ranges[i].second / VecOf<T>::N;
        }
        ownVec.n = ownScalar.n = (int)ranges.size();
        auto wins = E->processScalarRanges();
        winVec = WindowRanges{}; winScalar = WindowRanges{};
        if (!wins.empty()) {
            if (wins.size() > 3) { fprintf(stderr, "opt_b200: at most three process windows supported\n"); return -4; }
            for (size_t i = 0; i < wins.size(); ++i) {
                if (wins[i].first % VecOf<T>::N || wins[i].second % VecOf<T>::N) { fprintf(stderr, "opt_b200: process window not vector aligned\n"); return -3; }
                winScalar.lo[i] = wins[i].first; winScalar.len[i] = wins[i].second - wins[i].first; winScalar.total += winScalar.len[i];
                winVec.lo[i] = wins[i].first / VecOf<T>::N; winVec.len[i] = winScalar.len[i] / VecOf<T>::N; winVec.total += winVec.len[i];
            }
            winVec.n = winScalar.n = (int)wins.size();
        }
        buildZPush(segs, ranges, wins, upZ, dnZ);
        E->setPeerScratch((void*)(commBlock + 1), (rank > 0 && upperRows > 0) ? (void*)(peerBlocks[rank - 1] + 1) : nullptr,
                          (rank + 1 < world && lowerRows > 0) ? (void*)(peerBlocks[rank + 1] + 1) : nullptr);
        if (segs.size() > (size_t)kMaxHaloSegs || ranges.size() > 4) { fprintf(stderr, "opt_b200: at most %d halo segments / 4 owned ranges per rank are supported\n", kMaxHaloSegs); return -5; }
        if (loopExec) { cudaGraphExecDestroy(loopExec); loopExec = nullptr; }
        // collective decision about the one-kernel PCG loop: every rank posts whether it can run it, into every rank's block;
        // the first runLoop after this waits for all the answers (multiLoopAgreed)
        ++connectGen; multiLoopState = -1;
        const bool fits = persistentMode != 0 && !lm && !getenv("OPTB_NO_RESIDENT") && E->pcgLoopMultiFits();
        const unsigned long long word = ((unsigned long long)connectGen << 2) | (fits ? 2ull : 1ull);
        for (int r = 0; r < world; ++r) OPT_CUDA_CHECK(cudaMemcpy(&peerBlocks[r]->fitFlags[rank], &word, sizeof(word), cudaMemcpyDefault));
        return 0;
    }
    // Step2-push / Step3-wait form of the z halo: needs every boundary segment and every halo range aligned to the vector width
    template <class Segs, class Ranges, class Wins>
    void buildZPush(const Segs& segs, const Ranges& ranges, const Wins& wins, void* upZ, void* dnZ) {
        constexpr long NV = VecOf<T>::N;
        zPush = ZPush<T>{}; zWait = HaloWait{};
        if (getenv("OPTB_NO_ZPUSH") || segs.empty() || segs.size() > (size_t)kMaxHaloSegs || (!upZ && !dnZ)) return;
        ZPush<T> zp{}; HaloWait hw{};
        for (auto& sg : segs) {
            void* dst = sg.dir == 0 ? upZ : dnZ;
            if (!dst) continue;
            if (sg.srcOff % NV || sg.dstOff % NV || sg.len % NV) return;
            zp.lo[zp.n] = sg.srcOff / NV; zp.hi[zp.n] = (sg.srcOff + sg.len) / NV; zp.shift[zp.n] = (sg.dstOff - sg.srcOff) / NV; zp.dst[zp.n] = (T*)dst; ++zp.n;
        }
        zp.peerFlag[0] = haloZ.peerFlag[0]; zp.peerFlag[1] = haloZ.peerFlag[1];
        // halo entries = streamed entries (the windows, or the whole vectors) that are not owned
        std::vector<std::pair<long, long>> span;
        if (wins.empty()) span.push_back({ 0, n }); else for (auto& w : wins) span.push_back({ w.first, w.second });
        std::vector<std::pair<long, long>> own(ranges.begin(), ranges.end());
        std::sort(own.begin(), own.end());
        for (auto sp2 : span) {
            long at = sp2.first;
            for (auto& o : own) {
                if (o.second <= at || o.first >= sp2.second) continue;
                if (o.first > at) { if (hw.n >= 8 || at % NV || o.first % NV) return; hw.lo[hw.n] = at / NV; hw.len[hw.n] = (o.first - at) / NV; hw.total += hw.len[hw.n]; ++hw.n; }
                at = std::max(at, o.second);
            }
            if (at < sp2.second) { if (hw.n >= 8 || at % NV || sp2.second % NV) return; hw.lo[hw.n] = at / NV; hw.len[hw.n] = (sp2.second - at) / NV; hw.total += hw.len[hw.n]; ++hw.n; }
        }
        hw.flag[0] = upZ ? haloZ.myFlag[0] : nullptr; hw.flag[1] = dnZ ? haloZ.myFlag[1] : nullptr;
        if (zp.n == 0 || hw.n == 0) return;
        zPush = zp; zWait = hw;
    }
    unsigned connectGen = 0; int multiLoopState = -1;   // -1: not resolved yet, 0: per-step loop on every rank, 1: collective loop kernel on every rank
    bool multiLoopAgreed() {
        if (multiLoopState >= 0) return multiLoopState == 1;
        unsigned long long f[kMaxRanks];
        for (int spin = 0; ; ++spin) {   // a peer may still be inside its own OptB200_PlanConnect
            OPT_CUDA_CHECK(cudaMemcpy(f, commBlock->fitFlags, sizeof(f), cudaMemcpyDeviceToHost));
            bool all = true, yes = true;
            for (int r = 0; r < world; ++r) { if ((f[r] >> 2) != connectGen) all = false; else if ((f[r] & 3ull) != 2ull) yes = false; }
            if (all) { multiLoopState = yes ? 1 : 0; break; }
            if (spin > 60000) { fprintf(stderr, "opt_b200: rank %d: the other ranks never finished OptB200_PlanConnect\n", rank); exit(4); }
            struct timespec ts = { 0, 1000000 }; nanosleep(&ts, nullptr);
        }
        if (verbosity > 1) printf("opt_b200: rank %d: PCG loop as one collective kernel: %s\n", rank, multiLoopState ? "yes" : "no");
        return multiLoopState == 1;
    }
    CommBlock* peerBlocks[kMaxRanks] = {};
    unsigned long long* hostPeerTable[kMaxRanks] = {};
    void exchangeHalo(const HaloPlan& hp, const T* vec) {
        if (world <= 1 || hp.nSeg == 0 || (!hp.peerVec[0] && !hp.peerVec[1])) return;
        if (getenv("OPTB_SKIP_HALO")) return;   // timing experiment only (results are wrong)
        k_haloExchange<T><<<1, 1024, 0, stream>>>(&hp == &haloP ? dHaloP : (&hp == &haloZ ? dHaloZ : dHaloD), ws, vec); ++launches;
    }
    void readScalars() {
        resolveCost();
        if (world > 1) { k_finalizeScalars<<<1, 32, 0, stream>>>(ws, (1u << kRz0) | (1u << kRz1) | (1u << kPAp) | (1u << kCost) | (1u << kModelCost) | (1u << kQ)); ++launches; }
        OPT_CUDA_CHECK(cudaMemcpyAsync(hScalars, ws.scalars, sizeof(double) * kNumScalars, cudaMemcpyDeviceToHost, stream));
        OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    T computeCost() {   // :790-797
        E->cost(ctx()); ++launches;
        readScalars();
        return (T)hScalars[kCost];
    }

    void init(void** params) override {   // :956-1007
        E->bind(params);
        E->newSolve();
        enter();
        if (costPending || loopMsPending) { OPT_CUDA_CHECK(cudaStreamSynchronize(stream)); resolveCost(); }
        nIter = 0; finished = false; reported = false; launches = 0;
        if (timingOn) { OPT_CUDA_CHECK(cudaStreamSynchronize(stream)); dropTimedLaunches(); }   // a previous solve that was abandoned before it finished
        trust_region_radius = (T)sp.trust_region_radius; radius_decrease_factor = (T)sp.radius_decrease_factor;
        min_lm_diagonal = (T)sp.min_lm_diagonal; max_lm_diagonal = (T)sp.max_lm_diagonal;
        E->precompute(ctx());
        prevCost = computeCost();
        if (verbosity > 0) printf("initial cost=%f\n", (double)prevCost);
    }

    // ---- one PCG iteration, launched on `stream` (captured into the loop graph or issued directly)
    void launchStep2(int rzOld, int rzNew, T* p, bool pushZ = false, int zetaIter = -1) {
        const ZPush<T> zp = pushZ ? zPush : ZPush<T>{};
        const T zetaTol = (T)sp.q_tolerance;
        long nVec = n / VecOf<T>::N;
        if (lm) {
            if (E->usePreconditioner) k_pcgStep2<T, true, true><<<gridStep2, kBlock, 0, stream>>>(nVecWin(), winVec, ws, ownVec, rzOld, rzNew, delta, p, r, Ap, pre, b, z, zp, zetaIter, zetaTol);
            else k_pcgStep2<T, true, false><<<gridStep2, kBlock, 0, stream>>>(nVecWin(), winVec, ws, ownVec, rzOld, rzNew, delta, p, r, Ap, pre, b, z, zp, zetaIter, zetaTol);
        } else {
            if (E->usePreconditioner) launchPdl(k_pcgStep2<T, false, true>, gridStep2, kBlock, 0, stream, usePdl, nVecWin(), winVec, ws, ownVec, rzOld, rzNew, delta, p, r, Ap, pre, b, z, zp, zetaIter, zetaTol);
            else launchPdl(k_pcgStep2<T, false, false>, gridStep2, kBlock, 0, stream, usePdl, nVecWin(), winVec, ws, ownVec, rzOld, rzNew, delta, p, r, Ap, pre, b, z, zp, zetaIter, zetaTol);
        }
        ++launches;
    }
    void launchStep3(int rzNew, int rzOld, T* p, bool waitZ = false, int lIter = 0) {
        const HaloWait hw = waitZ ? zWait : HaloWait{};
        long nVec = n / VecOf<T>::N;
        if (lm) k_pcgStep3<T, true><<<gridStream, kBlock, 0, stream>>>(nVecWin(), winVec, ws, rzNew, rzOld, z, p, ownVec, hw, lIter);
        else launchPdl(k_pcgStep3<T, false>, gridStream, kBlock, 0, stream, usePdl, nVecWin(), winVec, ws, rzNew, rzOld, (const T*)z, p, ownVec, hw, lIter);
        ++launches;
    }
    // Emits iterations [0, L).  rz slot convention: iteration l reads rz[l&1] as "old", writes rz[(l+1)&1].
    // pendingStep3: when fusing, Step3 of iteration l is executed at the start of iteration l+1's JTJ kernel.
    void emitLoop(int L, bool fuse) {
        LaunchCtx c = ctx();
        const T* ctc = lm ? CtC : nullptr;
        bool pending = false;
        T *cur = p, *alt = p2;
        const bool kernelHalo = world > 1 && fuse && E->inKernelHalo();   // z / p0 halos move inside the J^T J kernels
        // single GPU: the loop as ONE launch with its state on chip, where the energy offers it (shape_from_shading: 3 launches per PCG
        // step otherwise, each bound by its own launch / drain latency at 640x480); what it does not cover is emitted kernel by kernel
        int l0 = 0;
        loopDidModelCost = false; loopDidUpdate = false;
        if (world == 1 && useOnChipLoop && !timingFine) {
            bool mc = false, up = false;
            const bool whole = maxLinIters < 0;   // (the debug hook leaves the PCG state for inspection: no model cost, no update)
            typename EnergyOps<T>::LoopHeadArgs ha{ SSq, min_lm_diagonal, max_lm_diagonal, headSaveSSq };
            l0 = E->pcgLoopOnChip(c, L, lm, lm ? sp.residual_reset_period : 0, (T)sp.q_tolerance, delta, r, z, cur, Ap, pre, b, ctc, (lm && whole) ? &mc : nullptr,
                                  lm ? prevX : nullptr, whole ? &up : nullptr, headInLoop ? &ha : nullptr);
            if (headInLoop && l0 != L) { fprintf(stderr, "opt_b200: the loop kernel was to run the iteration's head but covered %d of %d iterations\n", l0, L); exit(2); }
            if (l0 > 0) ++launches;
            if (l0 > 0 && l0 != L && (mc || up)) { fprintf(stderr, "opt_b200: pcgLoopOnChip covered %d of %d iterations but ran its tail\n", l0, L); exit(2); }
            loopDidModelCost = l0 == L && mc; loopDidUpdate = l0 == L && up;
        }
        for (int l = l0; l < L; ++l) {
            int rzOld = l & 1, rzNew = (l + 1) & 1;
            if (pending) {
                // Step3 of iteration l-1 (beta = rz[rzOld_now]/rz[rzNew_now^...]) fused into this Step1
                E->step3ApplyJTJ(c, z, cur, alt, Ap, ctc, /*rzNew of l-1*/ rzOld, /*rzOld of l-1*/ rzNew, kPAp, kernelHalo ? dHaloZ : nullptr);
                std::swap(cur, alt);
                ++launches; pending = false;
            } else {
                tBeginFine("PCGStep1");
                E->applyJTJ(c, cur, Ap, ctc, kPAp, (kernelHalo && l == 0) ? dHaloP : nullptr); ++launches;   // PCGStep1 (+_Graph)
                tEndFine();
            }
            bool reset = lm && ((l + 1) % sp.residual_reset_period) == 0;
            if (reset) {   // :1077-1083
                long nVec = n / VecOf<T>::N;
                k_pcgStep2FirstHalf<T><<<gridStream, kBlock, 0, stream>>>(nVecWin(), winVec, ws, rzOld, delta, cur); ++launches;
                E->applyJTJ(c, delta, Adelta, ctc, -1); ++launches;
                if (E->usePreconditioner) k_pcgStep2SecondHalf<T, true><<<gridStream, kBlock, 0, stream>>>(nVecWin(), winVec, ws, ownVec, rzNew, delta, r, Adelta, pre, b, z);
                else k_pcgStep2SecondHalf<T, false><<<gridStream, kBlock, 0, stream>>>(nVecWin(), winVec, ws, ownVec, rzNew, delta, r, Adelta, pre, b, z);
                ++launches;
            }
            const bool step3Fused = fuse && !lm && l + 1 < L;   // Step3 of this iteration runs inside the next Step1
            const bool zInStep2 = world > 1 && zPush.n > 0 && !reset && !step3Fused;   // Step2 pushes z's boundary entries, the standalone Step3 waits for them
            const bool zetaInStep2 = lm && world == 1 && !reset;   // single GPU: PCGStep2's finishing CTA takes the zeta decision
            if (!reset) {
                tBeginFine("PCGStep2");
                launchStep2(rzOld, rzNew, cur, zInStep2, zetaInStep2 ? l : -1);
                tEndFine();
            }
            if (!zInStep2 && !(kernelHalo && step3Fused)) exchangeHalo(haloZ, z);   // neighbours need z on my boundary rows for their p = z + beta p
            if (step3Fused) pending = true;   // defer Step3 into the next Step1
            else { tBeginFine("PCGStep3"); launchStep3(rzNew, rzOld, cur, zInStep2, l); tEndFine(); }
            if (lm && !zetaInStep2) { k_lmZeta<T><<<1, 1, 0, stream>>>(ws, l, (T)sp.q_tolerance); ++launches; }
        }
        pFinal = cur;
    }
    bool loopTakesHead(int L) { return lm && world == 1 && useOnChipLoop && !timingFine && L > 0 && E->pcgLoopTakesHead(L, lm); }
    bool loopFuses() {
        bool fuse = useFusion && !lm;
        if (fuse) fuse = E->step3ApplyJTJ(ctx(), nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, -2);   // capability probe, no launch
        return fuse;
    }
    void runLoop(int L) {
        bool fuse = loopFuses();
        OPT_CUDA_CHECK(cudaEventRecord(evLoop0, stream));
        if (persistentMode != 0 && !lm && L > 0 && (world == 1 || multiLoopAgreed())) {   // whole loop as one cooperative kernel where the energy offers it
            T* pf = nullptr;
            if (world > 1 && !pHaloFresh) { exchangeHalo(haloP, p); pHaloFresh = true; }   // the multi-rank form stages p's halo rows from memory
            if (E->pcgLoop(ctx(), L, persistentMode > 0, delta, r, z, p, p2, Ap, pre, &pf, world > 1 ? &haloZ : nullptr)) {
                pFinal = pf; ++launches; lastLoopPersistent = true;
                // the update of X on my halo rows must equal the owner's (no X exchange): the collective kernel writes delta
                // on owned rows only (the single-reduction form keeps no shadow copy of the neighbour's boundary row), so fetch the halo rows of delta
                if (world > 1) exchangeHalo(haloD, delta);
                OPT_CUDA_CHECK(cudaEventRecord(evLoop1, stream));
                return;
            }
        }
        lastLoopPersistent = false;
        if (useGraph && L > 0) {
            // The captured kernel nodes hold their arguments BY VALUE (weights, caller pointers, plan-owned buffers the energy may have
            // reallocated in bind()/pcgInit(), q_tolerance, the reset period): re-capture whenever any of that differs from what the
            // cached graph was built from.  Same topology => cudaGraphExecUpdate patches the node parameters in place (microseconds).
            unsigned long long key = E->captureKey();
            { const long v[] = { (long)L, (long)fuse, (long)sp.residual_reset_period, (long)world, (long)rank, (long)headInLoop, (long)headSaveSSq, (long)(maxLinIters >= 0) };
              key = EnergyOps<T>::hashBytes(v, sizeof(v), key); }
            key = EnergyOps<T>::hashValue(sp.q_tolerance, key); key = EnergyOps<T>::hashValue(min_lm_diagonal, key); key = EnergyOps<T>::hashValue(max_lm_diagonal, key);
            key = EnergyOps<T>::hashValue(winVec, key); key = EnergyOps<T>::hashValue(ownVec, key);
            if (!loopExec || loopExecKey != key) {
                cudaGraph_t g;
                int saved = launches;
                OPT_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                emitLoop(L, fuse);
                OPT_CUDA_CHECK(cudaStreamEndCapture(stream, &g));
                bool updated = false;
                if (loopExec) {
                    cudaGraphExecUpdateResultInfo info{};
                    if (cudaGraphExecUpdate(loopExec, g, &info) == cudaSuccess) updated = true;
                    else { cudaGetLastError(); cudaGraphExecDestroy(loopExec); loopExec = nullptr; }
                }
                if (!updated) OPT_CUDA_CHECK(cudaGraphInstantiate(&loopExec, g, 0));
                OPT_CUDA_CHECK(cudaGraphDestroy(g));
                loopExecKey = key; ++graphBuilds;
                loopLaunchCount = launches - saved; launches = saved; loopPFinal = pFinal;
            }
            pFinal = loopPFinal;
            OPT_CUDA_CHECK(cudaGraphLaunch(loopExec, stream));
            launches += loopLaunchCount;
        } else {
            emitLoop(L, fuse);
        }
        OPT_CUDA_CHECK(cudaEventRecord(evLoop1, stream));
    }
    int loopLaunchCount = 0; T* loopPFinal = nullptr; bool pHaloFresh = false;
    // ---- collectPerKernelTimingInfo (util.t:405-511, :820-842): event pairs aggregated by name in first-seen order and
    // printed when the solve ends (verbosity > 0), in the reference's table format.
    struct TimedLaunch { const char* name; cudaEvent_t a, b; };
    std::vector<TimedLaunch> timedLaunches;
    bool timingOn = false, timingFine = false;
    void tBeginFine(const char* name) { if (timingFine) tBegin(name); }
    void tEndFine() { if (timingFine) tEnd(); }
    void tBegin(const char* name) {
        if (!timingOn) return;
        TimedLaunch e{ name, nullptr, nullptr };
        OPT_CUDA_CHECK(cudaEventCreate(&e.a)); OPT_CUDA_CHECK(cudaEventCreate(&e.b));
        OPT_CUDA_CHECK(cudaEventRecord(e.a, stream));
        timedLaunches.push_back(e);
    }
    void dropTimedLaunches() { for (auto& e : timedLaunches) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); } timedLaunches.clear(); }
    void tEnd() { if (timingOn && !timedLaunches.empty()) OPT_CUDA_CHECK(cudaEventRecord(timedLaunches.back().b, stream)); }
    void finishSolve() {   // cleanup(pd), solverGPUGaussNewton.t:1009-1014
        if (verbosity > 0) printf("final cost=%f\n", (double)prevCost);
        if (!timingOn) return;
        OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
        std::vector<const char*> names; std::vector<float> total; std::vector<int> count;
        for (auto& e : timedLaunches) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) { cudaGetLastError(); cudaEventDestroy(e.a); cudaEventDestroy(e.b); continue; }   // region left early
            size_t i = 0; for (; i < names.size(); ++i) if (!strcmp(names[i], e.name)) break;
            if (i == names.size()) { names.push_back(e.name); total.push_back(0.f); count.push_back(0); }
            total[i] += ms; ++count[i];
            cudaEventDestroy(e.a); cudaEventDestroy(e.b);
        }
        timedLaunches.clear();
        if (verbosity > 0) {
            printf("--------------------------------------------------------\n");
            printf("        Kernel        |   Count  |   Total   | Average \n");
            printf("----------------------+----------+-----------+----------\n");
            for (size_t i = 0; i < names.size(); ++i) {
                printf("----------------------+----------+-----------+----------\n");
                printf(" %-20s |   %4d   | %8.3fms| %7.4fms\n", names[i], count[i], total[i], total[i] / count[i]);
            }
            printf("--------------------------------------------------------\n");
        }
    }
    // Whole inner loop as one persistent cooperative kernel where the energy offers it (image_warping GN).
    // -1 (default): only the state-on-chip variant, when the image fits; OPTB_PERSISTENT=1: also the L2-state variant;
    // OPTB_PERSISTENT=0: always the CUDA graph of per-step kernels.
    int persistentMode = 0;

    // ---- one nonlinear iteration as ONE graph launch (from the second iteration of a solve on, for energies whose entry points are pure
    // launches): PCGInit1, the LM diagonal kernels, the whole PCG loop, model cost, update, precompute, cost and the read-back of the scalars.
    // What changes from iteration to iteration (the trust-region radius) is read from device memory, so the graph is re-captured only when
    // the fingerprint of its by-value arguments changes (as for the loop graph).  The reference issues ~15 + 3 L launches and one blocking
    // copy per PCG step here; for shape_from_shading (L = 10) the launches outside the loop were 45 % of an iteration.
    cudaGraphExec_t iterExec = nullptr; unsigned long long iterKey = 0; int iterLaunchCount = 0; T* iterPFinal = nullptr;
    bool useIterGraph = true, lastLoopPersistent = false, useOnChipLoop = true, loopDidModelCost = false, loopDidUpdate = false, headInLoop = false, headSaveSSq = false;
    double* hRadius = nullptr;
    void uploadRadius() {
        if (!hRadius) OPT_CUDA_CHECK(cudaMallocHost((void**)&hRadius, sizeof(double)));
        *hRadius = (double)trust_region_radius;
        OPT_CUDA_CHECK(cudaMemcpyAsync(ws.scalars + kRadius, hRadius, sizeof(double), cudaMemcpyHostToDevice, stream));
    }
    void enqueueIterationForCapture(int L) {   // the launch sequence of step() without its timing hooks; nIter >= 1 (no PCGSaveSSq)
        LaunchCtx c = ctx();
        if (lm) OPT_CUDA_CHECK(cudaMemcpyAsync(ws.scalars + kRadius, hRadius, sizeof(double), cudaMemcpyHostToDevice, stream));
        headInLoop = loopTakesHead(L); headSaveSSq = false;
        const bool fusedHead = headInLoop || (lm && world == 1 && E->pcgInitLM(c, delta, r, pre, p, CtC, SSq, b, kRz0, min_lm_diagonal, max_lm_diagonal, false));
        if (headInLoop) {}
        else if (fusedHead) ++launches;
        else { E->pcgInit(c, delta, r, pre, p, kRz0); ++launches; }
        if (lm && !fusedHead) {
            E->computeCtC(c, CtC, T(1)); ++launches;
            k_lmFinalizeDiagonal<T><<<gridStream, kBlock, 0, stream>>>(winScalar.n ? winScalar.total : n, winScalar, ws, ownScalar, kRz0, CtC, SSq, pre, r, b, p, delta, min_lm_diagonal, max_lm_diagonal); ++launches;
            k_copyScalar<<<1, 1, 0, stream>>>(ws, kQPrev, kQ);
            k_setScalar<<<1, 1, 0, stream>>>(ws, kFlags, 0.0); launches += 2;
        }
        pHaloFresh = false;
        if (!(world > 1 && loopFuses() && E->inKernelHalo() && L > 0)) { exchangeHalo(haloP, p); pHaloFresh = true; }
        const bool fuse = loopFuses();
        OPT_CUDA_CHECK(cudaEventRecordWithFlags(evLoop0, stream, cudaEventRecordExternal));
        emitLoop(L, fuse);   // (the one-kernel loops are not captured: their reduction sequence numbers are launch arguments -- see launchIterationGraph)
        OPT_CUDA_CHECK(cudaEventRecordWithFlags(evLoop1, stream, cudaEventRecordExternal));
        if (lm && !loopDidModelCost) { E->modelCost(c, delta); ++launches; }
        if (!loopDidUpdate) for (auto& u : E->unknowns) {
            long cnt = u.elements * u.channels;
            if (lm) k_saveAndUpdate<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, prevX + u.offset, (T*)u.user, delta + u.offset);
            else k_axpyInto<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, (T*)u.user, delta + u.offset);
            ++launches;
        }
        E->precompute(c);
        E->cost(c); ++launches;
        if (world > 1) { k_finalizeScalars<<<1, 32, 0, stream>>>(ws, (1u << kRz0) | (1u << kRz1) | (1u << kPAp) | (1u << kCost) | (1u << kModelCost) | (1u << kQ)); ++launches; }
        OPT_CUDA_CHECK(cudaMemcpyAsync(hScalars, ws.scalars, sizeof(double) * kNumScalars, cudaMemcpyDeviceToHost, stream));
    }
    // returns false when this iteration has to take the launch-by-launch path
    bool launchIterationGraph(int L) {
        if (!useIterGraph || !useGraph || nIter == 0 || timingOn || maxLinIters >= 0 || L <= 0 || !E->capturable()) return false;
        if (lastLoopPersistent) return false;   // the whole-loop cooperative kernels take their reduction sequence base as a launch argument (a replayed
                                                // graph would repeat it); they already are one launch per iteration
        if (world > 1 && multiLoopState < 0) return false;          // the collective decision is taken (with host polling) in the first iteration
        if (!hRadius) OPT_CUDA_CHECK(cudaMallocHost((void**)&hRadius, sizeof(double)));
        const bool fuse = loopFuses();
        unsigned long long key = E->captureKey();
        { const long v[] = { (long)L, (long)fuse, (long)lm, (long)sp.residual_reset_period, (long)world, (long)rank, (long)persistentMode, (long)multiLoopState, (long)n }; key = EnergyOps<T>::hashBytes(v, sizeof(v), key); }
        key = EnergyOps<T>::hashValue(sp.q_tolerance, key); key = EnergyOps<T>::hashValue(min_lm_diagonal, key); key = EnergyOps<T>::hashValue(max_lm_diagonal, key);
        key = EnergyOps<T>::hashValue(winVec, key); key = EnergyOps<T>::hashValue(ownVec, key);
        if (!iterExec || iterKey != key) {
            cudaGraph_t g;
            const int saved = launches;
            OPT_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            enqueueIterationForCapture(L);
            OPT_CUDA_CHECK(cudaStreamEndCapture(stream, &g));
            bool updated = false;
            if (iterExec) {
                cudaGraphExecUpdateResultInfo info{};
                if (cudaGraphExecUpdate(iterExec, g, &info) == cudaSuccess) updated = true;
                else { cudaGetLastError(); cudaGraphExecDestroy(iterExec); iterExec = nullptr; }
            }
            if (!updated) OPT_CUDA_CHECK(cudaGraphInstantiate(&iterExec, g, 0));
            OPT_CUDA_CHECK(cudaGraphDestroy(g));
            iterKey = key; iterLaunchCount = launches - saved; launches = saved; iterPFinal = pFinal; ++graphBuilds;
        }
        if (lm) *hRadius = (double)trust_region_radius;
        pFinal = iterPFinal;
        OPT_CUDA_CHECK(cudaGraphLaunch(iterExec, stream));
        launches += iterLaunchCount;
        return true;
    }

    int step(void** params) override {   // :1016-1177
        E->bind(params);
        if (finished || !(nIter < sp.nIterations)) {
            if (costPending || loopMsPending) { OPT_CUDA_CHECK(cudaStreamSynchronize(stream)); resolveCost(); }   // the solve ends with all of its work complete, like the reference
            if (!reported) { finishSolve(); reported = true; }
            finished = true; return 0;
        }
        enter();
        launches = 0;
        LaunchCtx c = ctx();
        {
            int Lg = sp.lIterations;
            if (!lm && verbosity == 0 && asyncSteps) resolveCost();     // the previous step's value, before its pinned slot is reused
            if (launchIterationGraph(Lg)) {
                if (!lm && verbosity == 0 && asyncSteps) {              // GN: the cost stays behind an event (see computeCostAsync)
                    OPT_CUDA_CHECK(cudaEventRecord(evCost, stream)); costPending = true;
                    OPT_CUDA_CHECK(cudaEventRecord(evExit, stream)); OPT_CUDA_CHECK(cudaStreamWaitEvent(userStream, evExit, 0));
                    loopMsPending = true; ++nIter;
                    return 1;
                }
                OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
                OPT_CUDA_CHECK(cudaEventElapsedTime(&loopMs, evLoop0, evLoop1)); loopMsPending = false;
                return decide((T)hScalars[kCost], c);
            }
        }
        if (lm) uploadRadius();
        tBegin("PCGInit1");
        // one launch for PCGInit1 and the LM head (:1038-1053) where the energy offers it
        int Lh = sp.lIterations; if (maxLinIters >= 0 && maxLinIters < Lh) Lh = maxLinIters;
        headInLoop = loopTakesHead(Lh); headSaveSSq = nIter == 0;        // the one-launch loop may run the head on its tiles itself
        const bool fusedHead = headInLoop || (lm && world == 1 && E->pcgInitLM(c, delta, r, pre, p, CtC, SSq, b, kRz0, min_lm_diagonal, max_lm_diagonal, nIter == 0));
        if (!fusedHead) E->pcgInit(c, delta, r, pre, p, kRz0);          // PCGInit1 (+_Graph,+_Finish)
        if (!headInLoop) ++launches;
        tEnd();
        if (lm && !fusedHead) {   // :1038-1053
            if (nIter == 0) { k_copy<T><<<gridStream, kBlock, 0, stream>>>(n, SSq, pre); ++launches; }   // PCGSaveSSq
            E->computeCtC(c, CtC, T(1)); ++launches;
            k_lmFinalizeDiagonal<T><<<gridStream, kBlock, 0, stream>>>(winScalar.n ? winScalar.total : n, winScalar, ws, ownScalar, kRz0, CtC, SSq, pre, r, b, p, delta, min_lm_diagonal, max_lm_diagonal); ++launches;
            k_copyScalar<<<1, 1, 0, stream>>>(ws, kQPrev, kQ);          // Q0 = fetchQ
            k_setScalar<<<1, 1, 0, stream>>>(ws, kFlags, 0.0); launches += 2;
        }
        int L = sp.lIterations; if (maxLinIters >= 0 && maxLinIters < L) L = maxLinIters;
        pHaloFresh = false;
        if (!(world > 1 && loopFuses() && E->inKernelHalo() && L > 0)) { exchangeHalo(haloP, p); pHaloFresh = true; }   // initial direction on the halo rows
        if (!timingFine) tBegin("PCG loop");
        runLoop(L);
        if (!timingFine) tEnd();
        if (maxLinIters >= 0) { OPT_CUDA_CHECK(cudaStreamSynchronize(stream)); return 1; }   // debug hook: PCG state left for inspection
        if (lm && !loopDidModelCost) { E->modelCost(c, delta); ++launches; }   // :1108-1113 (the one-launch loop may have left it already)
        tBegin("PCGLinearUpdate");
        if (!loopDidUpdate) for (auto& u : E->unknowns) {   // LM: savePreviousUnknowns, then PCGLinearUpdate :552-557 (delta is 0 on excluded entries)
            long cnt = u.elements * u.channels;
            if (lm) k_saveAndUpdate<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, prevX + u.offset, (T*)u.user, delta + u.offset);
            else k_axpyInto<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, (T*)u.user, delta + u.offset);
            ++launches;
        }
        tEnd();
        E->precompute(c);
        if (!lm && verbosity == 0 && asyncSteps) {   // GN: nothing below needs the cost on the host (the timing table, when asked for, is
            resolveCost();                           // assembled from event pairs when the solve ends)
            tBegin("computeCost");
            computeCostAsync();
            tEnd();
            loopMsPending = true;
            ++nIter;
            return 1;
        }
        tBegin("computeCost");
        T newCost = computeCost();
        tEnd();    // syncs; hScalars also carries the model cost and the loop timing is complete
        OPT_CUDA_CHECK(cudaEventElapsedTime(&loopMs, evLoop0, evLoop1)); loopMsPending = false;
        return decide(newCost, c);
    }
    // ---- LM with the decision on the device (opt-in: OPTB_DEVICE_LM=1; single GPU; energies that answer deviceLmCapable()).
    // Opt_ProblemSolve calls this after Init.  The first nonlinear iteration runs through step() (PCGSaveSSq, and the host needs one
    // decision to seed the state); the nonlinear state (radius, previous cost, decrease factor) then moves into the scalar block and ALL
    // remaining iterations are enqueued back to back, each one graph launch:
    //     loop kernel (head, PCG loop, model cost, save + update)   -- returns at once when scalars[kLmGate] != 0
    //     precompute + cost                                         -- likewise
    //     k_lmDecide          accept / reject / radius / end of solve (lm_decide.cuh, the arithmetic of decide() below)
    //     k_copyIfRejected    revertUpdate
    //     precompute + cost   only after a rejection
    // and the host waits once per chunk of iterations (8), for the state.  An iteration enqueued behind the one that ended the solve
    // costs five empty launches.  Opt_ProblemStep keeps the host-side decision (it has to return it).
    cudaGraphExec_t devIterExec = nullptr; unsigned long long devIterKey = 0; int devIterLaunchCount = 0; T* devIterPFinal = nullptr;
    double* hLmState = nullptr; int deviceIterationsRun = 0;
    void runDeviceIterations(void** params) override {
        deviceIterationsRun = 0;
        const char* want = getenv("OPTB_DEVICE_LM");
        if (!want || atoi(want) == 0 || !lm || world != 1 || finished) return;
        if (!useIterGraph || !useGraph || timingOn || maxLinIters >= 0 || verbosity > 0 || !E->capturable()) return;
        const int L = sp.lIterations;
        if (nIter == 0) { if (!(nIter < sp.nIterations) || !step(params)) return; }
        if (finished || !(nIter < sp.nIterations)) return;
        E->bind(params);
        if (!E->deviceLmCapable() || !loopTakesHead(L)) return;
        enter();
        const LmLimits<T> lim{ (T)sp.min_relative_decrease, (T)sp.function_tolerance, (T)sp.max_trust_region_radius, (T)sp.min_trust_region_radius };
        unsigned long long key = E->captureKey();
        { const long v[] = { (long)L, (long)sp.residual_reset_period, (long)n }; key = EnergyOps<T>::hashBytes(v, sizeof(v), key); }
        key = EnergyOps<T>::hashValue(sp.q_tolerance, key); key = EnergyOps<T>::hashValue(min_lm_diagonal, key); key = EnergyOps<T>::hashValue(max_lm_diagonal, key);
        key = EnergyOps<T>::hashValue(lim, key);
        if (!devIterExec || devIterKey != key) {
            cudaGraph_t g;
            const int saved = launches;
            OPT_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            ctxGate = 1;
            headInLoop = true; headSaveSSq = false;
            emitLoop(L, false);
            const bool whole = loopDidModelCost && loopDidUpdate;
            E->precompute(ctx()); ++launches;
            k_lmDecide<T><<<1, 1, 0, stream>>>(ws, lim); ++launches;
            for (auto& u : E->unknowns) { long cnt = u.elements * u.channels; k_copyIfRejected<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, (T*)u.user, prevX + u.offset, ws); ++launches; }
            ctxGate = 2;
            E->precompute(ctx()); ++launches;
            ctxGate = 0;
            OPT_CUDA_CHECK(cudaStreamEndCapture(stream, &g));
            const int count = launches - saved; launches = saved;
            if (!whole) { OPT_CUDA_CHECK(cudaGraphDestroy(g)); return; }   // the loop kernel left its tail to separate launches: host path
            bool updated = false;
            if (devIterExec) {
                cudaGraphExecUpdateResultInfo info{};
                if (cudaGraphExecUpdate(devIterExec, g, &info) == cudaSuccess) updated = true;
                else { cudaGetLastError(); cudaGraphExecDestroy(devIterExec); devIterExec = nullptr; }
            }
            if (!updated) OPT_CUDA_CHECK(cudaGraphInstantiate(&devIterExec, g, 0));
            OPT_CUDA_CHECK(cudaGraphDestroy(g));
            devIterKey = key; devIterLaunchCount = count; devIterPFinal = pFinal; ++graphBuilds;
        }
        if (!hLmState) OPT_CUDA_CHECK(cudaMallocHost((void**)&hLmState, sizeof(double) * 6));
        static_assert(kLmPrevCost == kRadius + 1 && kLmFactor == kRadius + 2 && kLmGate == kRadius + 3 && kLmReject == kRadius + 4 && kLmIters == kRadius + 5 && kLmIters + 1 == kNumScalars,
                      "the nonlinear state is uploaded as one block behind kRadius");
        hLmState[0] = (double)trust_region_radius; hLmState[1] = (double)prevCost; hLmState[2] = (double)radius_decrease_factor;
        hLmState[3] = hLmState[4] = hLmState[5] = 0.0;
        OPT_CUDA_CHECK(cudaMemcpyAsync(ws.scalars + kRadius, hLmState, sizeof(double) * 6, cudaMemcpyHostToDevice, stream));
        // enqueued in chunks (default 8 iterations; OPTB_DEVICE_LM=n > 1 sets n): one wait per chunk instead of one per iteration, and a solve
        // that ends early leaves at most chunk - 1 gated (empty) iterations behind it
        const int chunk = atoi(want) > 1 ? atoi(want) : 8;
        for (int left = sp.nIterations - nIter; left > 0; left -= chunk) {
            for (int i = 0; i < chunk && i < left; ++i) OPT_CUDA_CHECK(cudaGraphLaunch(devIterExec, stream));
            OPT_CUDA_CHECK(cudaMemcpyAsync(hScalars, ws.scalars, sizeof(double) * kNumScalars, cudaMemcpyDeviceToHost, stream));
            OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
            if (hScalars[kLmGate] != 0.0) break;
        }
        pFinal = devIterPFinal; launches = devIterLaunchCount; loopMs = 0;
        prevCost = (T)hScalars[kLmPrevCost]; trust_region_radius = (T)hScalars[kRadius]; radius_decrease_factor = (T)hScalars[kLmFactor];
        deviceIterationsRun = (int)hScalars[kLmIters];
        nIter += deviceIterationsRun;
        if (hScalars[kLmGate] != 0.0) { finished = true; if (!reported) { finishSolve(); reported = true; } }
    }
    // the host side of the nonlinear iteration: accept / reject / trust-region update (:1119-1157), with the cost and model cost already read back
    int decide(T newCost, const LaunchCtx& c) {
        T model_cost_change = 0;
        int ret = 1;
        if (lm) {   // :1119-1157
            T model_cost = (T)hScalars[kModelCost];
            model_cost_change = prevCost - model_cost;
            T cost_change = prevCost - newCost;
            T relative_decrease = cost_change / model_cost_change;
            if (verbosity > 0) printf(" cost=%f model_cost=%f model_cost_change=%f newCost=%f\n", (double)prevCost, (double)model_cost, (double)model_cost_change, (double)newCost);
            if (cost_change >= 0 && relative_decrease > (T)sp.min_relative_decrease) {
                T absolute_function_tolerance = prevCost * (T)sp.function_tolerance;
                if (cost_change <= absolute_function_tolerance) {
                    if (verbosity > 0) printf("\nFunction tolerance reached, exiting\n");
                    finished = true; if (!reported) { finishSolve(); reported = true; } return 0;
                }
                T step_quality = relative_decrease;
                double min_factor = 1.0 / 3.0;
                double tmp_factor = 1.0 - pow(2.0 * (double)step_quality - 1.0, 3.0);
                trust_region_radius = (T)((double)trust_region_radius / fmax(min_factor, tmp_factor));
                trust_region_radius = (T)fmin((double)trust_region_radius, (double)(T)sp.max_trust_region_radius);
                radius_decrease_factor = T(2.0);
                prevCost = newCost;
            } else {
                for (auto& u : E->unknowns) { long cnt = u.elements * u.channels; k_copy<T><<<gridFor(cnt), kBlock, 0, stream>>>(cnt, (T*)u.user, prevX + u.offset); }   // revertUpdate
                trust_region_radius = trust_region_radius / radius_decrease_factor;
                if (verbosity > 0) printf(" trust_region_radius=%f \nREVERT\n", (double)trust_region_radius);
                radius_decrease_factor = T(2.0) * radius_decrease_factor;
                if (trust_region_radius <= (T)sp.min_trust_region_radius) {
                    if (verbosity > 0) printf("\nTrust_region_radius is less than the min, exiting\n");
                    OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
                    finished = true; if (!reported) { finishSolve(); reported = true; } return 0;
                }
                E->precompute(c);
                OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
            }
        } else {
            if (verbosity > 0) printf("cost: %f -> %f\n", (double)prevCost, (double)newCost);
            prevCost = newCost;
        }
        ++nIter;
        return ret;
    }
    int gridFor(long cnt) const { long g = (cnt + kBlock - 1) / kBlock; if (g > (long)numSMs * 8) g = (long)numSMs * 8; return g < 1 ? 1 : (int)g; }

    // ---- extension entry points used by the parity tests (include/OptB200.h)
    void evalJTF(void** params, double* costOut, void* negJtF, void* preInv, double* rDotP) override {
        E->bind(params); enter();
        LaunchCtx c = ctx();
        E->precompute(c);
        E->cost(c);
        E->pcgInit(c, delta, r, pre, p, kRz0);
        readScalars();
        if (costOut) *costOut = hScalars[kCost];
        if (rDotP) *rDotP = hScalars[kRz0];
        if (negJtF) OPT_CUDA_CHECK(cudaMemcpy(negJtF, r, n * sizeof(T), cudaMemcpyDeviceToHost));
        if (preInv) OPT_CUDA_CHECK(cudaMemcpy(preInv, pre, n * sizeof(T), cudaMemcpyDeviceToHost));
    }
    void applyJTJ(void** params, const void* x, void* out, double* xDotOut) override {
        E->bind(params); enter();
        LaunchCtx c = ctx();
        E->precompute(c);
        E->pcgInit(c, delta, r, pre, p, kRz0);   // refresh the linearisation
        OPT_CUDA_CHECK(cudaMemcpyAsync(p, x, n * sizeof(T), cudaMemcpyHostToDevice, stream));
        E->applyJTJ(c, p, Ap, nullptr, kPAp);
        readScalars();
        if (xDotOut) *xDotOut = hScalars[kPAp];
        OPT_CUDA_CHECK(cudaMemcpy(out, Ap, n * sizeof(T), cudaMemcpyDeviceToHost));
    }
    // Event-timed replay of each inner-loop kernel on the plan's stream (average over `reps` back-to-back
    // launches after 3 warm-ups).  Destroys the PCG state; callers re-Init afterwards.
    int profileKernels(void** params, int reps, int maxOut, const char** names, float* avgMs) override {
        E->bind(params); enter();
        LaunchCtx c = ctx();
        E->precompute(c);
        E->pcgInit(c, delta, r, pre, p, kRz0);
        const T* ctc = lm ? CtC : nullptr;
        if (lm) { E->computeCtC(c, CtC, T(1) / (T)sp.trust_region_radius); OPT_CUDA_CHECK(cudaMemsetAsync(b, 0, n * sizeof(T), stream)); }
        bool fused = E->step3ApplyJTJ(c, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, -2);
        struct K { const char* name; int id; } ks[] = { { "PCGInit1", 0 }, { "PCGStep1", 1 }, { "PCGStep3+PCGStep1(fused)", 2 }, { "PCGStep2", 3 }, { "PCGStep3", 4 }, { "computeCost", 5 } };
        int k = 0;
        cudaEvent_t e0, e1; OPT_CUDA_CHECK(cudaEventCreate(&e0)); OPT_CUDA_CHECK(cudaEventCreate(&e1));
        for (auto& kk : ks) {
            if (kk.id == 2 && !fused) continue;
            if (k >= maxOut) break;
            auto launch = [&]() {
                switch (kk.id) {
                    case 0: E->pcgInit(c, delta, r, pre, p, kRz0); break;
                    case 1: E->applyJTJ(c, p, Ap, ctc, kPAp); break;
                    case 2: E->step3ApplyJTJ(c, z, p, p2, Ap, ctc, kRz1, kRz0, kPAp); break;
                    case 3: launchStep2(kRz0, kRz1, p); break;
                    case 4: launchStep3(kRz1, kRz0, p); break;
                    case 5: E->cost(c); break;
                }
            };
            for (int w = 0; w < 3; ++w) launch();
            OPT_CUDA_CHECK(cudaEventRecord(e0, stream));
            for (int i = 0; i < reps; ++i) launch();
            OPT_CUDA_CHECK(cudaEventRecord(e1, stream));
            OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
            float ms = 0; OPT_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
            names[k] = kk.name; avgMs[k] = ms / reps; ++k;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return k;
    }
    int readVector(const char* name, void* hostOut) override {
        std::string s(name); const T* v = nullptr;
        if (s == "delta") v = delta; else if (s == "r") v = r; else if (s == "p") v = pFinal; else if (s == "z") v = z;
        else if (s == "Ap_X") v = Ap; else if (s == "preconditioner") v = pre; else if (s == "CtC") v = CtC; else if (s == "b") v = b;
        if (!v) return -1;
        OPT_CUDA_CHECK(cudaStreamSynchronize(stream));
        OPT_CUDA_CHECK(cudaMemcpy(hostOut, v, n * sizeof(T), cudaMemcpyDeviceToHost));
        return 0;
    }
    double readScalar(const char* name) override {
        std::string s(name);
        readScalars();
        if (s == "scanAlphaDenominator") return hScalars[kPAp];
        if (s == "rz0") return hScalars[kRz0]; if (s == "rz1") return hScalars[kRz1];
        if (s == "q") return hScalars[kQ]; if (s == "flags") return hScalars[kFlags];
        if (s == "trust_region_radius") return (double)trust_region_radius;
        if (s == "cost") return hScalars[kCost]; if (s == "modelCost") return hScalars[kModelCost];
        if (s == "deviceIterations") return (double)deviceIterationsRun; if (s == "nIter") return (double)nIter;
        return NAN;
    }
};

template <class T>
static EnergyOps<T>* makeEnergy(const std::string& energy, const unsigned* dims) {
    // OPTB_GENERIC=1 routes every energy (image_warping included) through the generic cached-Jacobian engine
    if (energy == "image_warping" && !getenv("OPTB_GENERIC")) return makeImageWarping<T>(dims);
    if (energy == "arap_mesh_deformation" && !getenv("OPTB_GENERIC")) return makeArapMesh<T>(dims);
    if (energy == "robust_nonrigid_alignment" && !getenv("OPTB_GENERIC")) return makeRna<T>(dims);
    if (energy == "shape_from_shading" && !getenv("OPTB_GENERIC")) return makeSfs<T>(dims);
    return makeGenericEnergy<T>(energy, dims);
}

PlanBase* makePlan(const std::string& energy, bool lm, bool doublePrecision, const unsigned* dims, int verbosity, int collectTiming) {
    if (doublePrecision) {
        EnergyOps<double>* e = makeEnergy<double>(energy, dims);
        return e ? new Plan<double>(e, lm, verbosity, collectTiming) : nullptr;
    }
    EnergyOps<float>* e = makeEnergy<float>(energy, dims);
    return e ? new Plan<float>(e, lm, verbosity, collectTiming) : nullptr;
}

PlanBase* makePlanDynamic(const fe::CompiledSpec& cs, const std::string& name, bool lm, bool doublePrecision, const unsigned* dims, int verbosity,
                          int collectTiming, std::string& error) {
    if (doublePrecision) {
        EnergyOps<double>* e = makeDynamicEnergy<double>(cs, dims, name, error);
        return e ? new Plan<double>(e, lm, verbosity, collectTiming) : nullptr;
    }
    EnergyOps<float>* e = makeDynamicEnergy<float>(cs, dims, name, error);
    return e ? new Plan<float>(e, lm, verbosity, collectTiming) : nullptr;
}

}  // namespace optb
