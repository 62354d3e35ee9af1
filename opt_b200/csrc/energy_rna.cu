// energy_rna.cu -- hand-written kernel set for robust_nonrigid_alignment (BASELINE config "robust_nonrigid_alignment
// (graph+image, robust LM) fp64").
//
// All per-vertex math lives in rna_device.cuh (host/device code, verified on the CPU against the oracle by
// tests/test_rna_host_parity.py); the kernels here are one-thread-per-vertex grid-stride loops around those functions, the
// vertex-CSR construction (stable cub radix sorts of the caller's index arrays, as in energy_arap.cu) and the deterministic
// reductions of common.cuh.  Replaces, for this energy, the generated function map of the reference (o.t:2029-2460) and its
// per-edge atomicAdd scatter (solverGPUGaussNewton.t:687-706).
// Default for this energy (OPTB_GENERIC=1 selects the cached-Jacobian engine instead); parity: tests/test_gpu_generic_energies.py.
#include <algorithm>
#include <cub/cub.cuh>
#include "solver.h"
#include "rna_device.cuh"

namespace optb {

__device__ __forceinline__ float rnaGuardedInvert(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }     // CERES variant, solverGPUGaussNewton.t:323-332
__device__ __forceinline__ double rnaGuardedInvert(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }

template <class T>
__global__ void __launch_bounds__(kBlock) k_rnaPack(RnaData<T> D) {
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) D.pack[v] = rnaMakePack(D, v);
}
template <class T> __device__ __forceinline__ void rnaStore(T* vec, const RnaData<T>& D, int v, const RnaOut<T>& q) {
    const long o = 3L * v, a = D.offA + 3L * v;
    vec[o] = q.O.x; vec[o + 1] = q.O.y; vec[o + 2] = q.O.z; vec[a] = q.A.x; vec[a + 1] = q.A.y; vec[a + 2] = q.A.z; vec[D.offG + v] = q.g;
}
// PCGInit1 (+_Graph,+_Finish), :361-419, :687-692: r = -J^T F, pre = guardedInvert(diag J^T J), p = pre r, delta = 0, r.p
template <class T>
__global__ void __launch_bounds__(kBlock) k_rnaInit(RnaData<T> D, ReduceWorkspace ws, int rzSlot, int usePre, T* __restrict__ delta, T* __restrict__ r,
                                                    T* __restrict__ pre, T* __restrict__ p) {
    T acc = T(0);
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        RnaOut<T> rr, pr, pp, zero;
        zero.O = RnaV3<T>{ T(0), T(0), T(0) }; zero.A = zero.O; zero.g = T(0);
        if (v < D.lo || v >= D.hi) { rnaStore(delta, D, v, zero); rnaStore(r, D, v, zero); rnaStore(pre, D, v, zero); continue; }   // halo vertex: its p arrives from the owner
        const RnaOut<T> g = rnaGather<T, 1>(D, v, nullptr), d = rnaGather<T, 2>(D, v, nullptr);
        rr.O = RnaV3<T>{ -g.O.x, -g.O.y, -g.O.z }; rr.A = RnaV3<T>{ -g.A.x, -g.A.y, -g.A.z }; rr.g = -g.g;
        // graph problems: PCGInit1_Finish order -- invert first, then the UsePreconditioner(false) override (:406-410)
        pr.O = RnaV3<T>{ rnaGuardedInvert(d.O.x), rnaGuardedInvert(d.O.y), rnaGuardedInvert(d.O.z) };
        pr.A = RnaV3<T>{ rnaGuardedInvert(d.A.x), rnaGuardedInvert(d.A.y), rnaGuardedInvert(d.A.z) };
        pr.g = rnaGuardedInvert(d.g);
        if (!usePre) { pr.O = RnaV3<T>{ T(1), T(1), T(1) }; pr.A = pr.O; pr.g = T(1); }
        pp.O = RnaV3<T>{ pr.O.x * rr.O.x, pr.O.y * rr.O.y, pr.O.z * rr.O.z }; pp.A = RnaV3<T>{ pr.A.x * rr.A.x, pr.A.y * rr.A.y, pr.A.z * rr.A.z }; pp.g = pr.g * rr.g;
        acc += rnaDot(rr.O, pp.O) + rnaDot(rr.A, pp.A) + rr.g * pp.g;
        rnaStore(delta, D, v, zero); rnaStore(r, D, v, rr); rnaStore(pre, D, v, pr); rnaStore(p, D, v, pp);
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}
// PCGStep1 (+_Graph), :421-434, :694-706: out = (J^T J [+CtC]) x, x.out
template <class T, bool LM>
__global__ void __launch_bounds__(kBlock) k_rnaApplyJTJ(RnaData<T> D, ReduceWorkspace ws, int dotSlot, const T* __restrict__ x, T* __restrict__ out, const T* __restrict__ CtC) {
    T acc = T(0);
    for (int v = D.wlo + blockIdx.x * kBlock + threadIdx.x; v < D.whi; v += gridDim.x * kBlock) {
        if (v < D.lo || v >= D.hi) { RnaOut<T> zero; zero.O = RnaV3<T>{ T(0), T(0), T(0) }; zero.A = zero.O; zero.g = T(0); rnaStore(out, D, v, zero); continue; }
        RnaOut<T> o = rnaGather<T, 0>(D, v, x);
        const long a0 = 3L * v, a1 = D.offA + 3L * v, a2 = D.offG + v;
        const RnaV3<T> xO{ x[a0], x[a0 + 1], x[a0 + 2] }, xA{ x[a1], x[a1 + 1], x[a1 + 2] }; const T xg = x[a2];
        if (LM) {
            o.O.x += CtC[a0] * xO.x; o.O.y += CtC[a0 + 1] * xO.y; o.O.z += CtC[a0 + 2] * xO.z;
            o.A.x += CtC[a1] * xA.x; o.A.y += CtC[a1 + 1] * xA.y; o.A.z += CtC[a1 + 2] * xA.z;
            o.g += CtC[a2] * xg;
        }
        acc += rnaDot(xO, o.O) + rnaDot(xA, o.A) + xg * o.g;
        rnaStore(out, D, v, o);
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}
template <class T, bool MODEL>
__global__ void __launch_bounds__(kBlock) k_rnaCost(RnaData<T> D, ReduceWorkspace ws, int slot, const T* __restrict__ delta) {
    T acc = T(0);
    for (int v = D.lo + blockIdx.x * kBlock + threadIdx.x; v < D.hi; v += gridDim.x * kBlock) acc += rnaCostAt<T, MODEL>(D, v, delta);
    gridReduceTo<kBlock>((double)acc, ws, slot);
}
// PCGComputeCtC (+_Graph), :616-622, :739-744: CtC = diag(J^T J) / radius
template <class T>
__global__ void __launch_bounds__(kBlock) k_rnaCtC(RnaData<T> D, T* __restrict__ CtC, T invRadius) {
    for (int v = D.lo + blockIdx.x * kBlock + threadIdx.x; v < D.hi; v += gridDim.x * kBlock) {
        RnaOut<T> d = rnaGather<T, 2>(D, v, nullptr);
        d.O.x *= invRadius; d.O.y *= invRadius; d.O.z *= invRadius; d.A.x *= invRadius; d.A.y *= invRadius; d.A.z *= invRadius; d.g *= invRadius;
        rnaStore(CtC, D, v, d);
    }
}
// graph symmetry test (as energy_arap.cu): every out-neighbour b of v must list v among ITS out-neighbours with the same multiplicity
static __global__ void k_rnaSymmetric(int N, const int* __restrict__ outStart, const int* __restrict__ outNbr, const int* __restrict__ inStart, int* __restrict__ notSymmetric) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < N; v += gridDim.x * blockDim.x) {
        bool ok = (outStart[v + 1] - outStart[v]) == (inStart[v + 1] - inStart[v]);
        for (int k = outStart[v]; ok && k < outStart[v + 1]; ++k) {
            const int b = outNbr[k];
            int cnt = 0, mine = 0;
            for (int q = outStart[b]; q < outStart[b + 1]; ++q) cnt += outNbr[q] == v;
            for (int q = outStart[v]; q < outStart[v + 1]; ++q) mine += outNbr[q] == b;
            ok = cnt == mine;
        }
        if (!ok) *notSymmetric = 1;
    }
}
static __global__ void k_rnaCsrStart(int nE, const int* __restrict__ sortedKeys, int N, int* __restrict__ start) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= nE; t += gridDim.x * blockDim.x) {
        const int k = t < nE ? sortedKeys[t] : N, kp = t > 0 ? sortedKeys[t - 1] : -1;
        for (int u = kp + 1; u <= k && u <= N; ++u) start[u] = t;
    }
}

template <class T>
struct RnaOps : EnergyOps<T> {
    RnaData<T> D{};
    const int *v0 = nullptr, *v1 = nullptr; int nEdges = 0;
    int *outStart = nullptr, *outNbr = nullptr, *inStart = nullptr, *inNbr = nullptr, *keysTmp = nullptr;
    void* cubTemp = nullptr; size_t cubBytes = 0; int capE = 0;
    bool csrValid = false, packValid = false;
    int sms = 148;
    RnaOps(const unsigned* dims) {
        D.N = (int)dims[0]; D.lo = D.wlo = 0; D.hi = D.whi = D.N;
        this->unknowns.resize(3);
        this->unknowns[0].elements = D.N; this->unknowns[0].channels = 3;   // Offset
        this->unknowns[1].elements = D.N; this->unknowns[1].channels = 3;   // Angle
        this->unknowns[2].elements = D.N; this->unknowns[2].channels = 1;   // RobustWeights
        this->layoutUnknowns();
        D.offA = this->unknowns[1].offset; D.offG = this->unknowns[2].offset;
        D.wc = T(0.1);
        this->usePreconditioner = true; this->usesGraph = true;
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        OPT_CUDA_CHECK(cudaMalloc((void**)&D.pack, sizeof(RnaPack<T>) * (size_t)D.N));
        OPT_CUDA_CHECK(cudaMalloc((void**)&outStart, sizeof(int) * (D.N + 2))); OPT_CUDA_CHECK(cudaMalloc((void**)&inStart, sizeof(int) * (D.N + 2)));
    }
    ~RnaOps() override { for (void* q : { (void*)D.pack, (void*)outStart, (void*)outNbr, (void*)inStart, (void*)inNbr, (void*)keysTmp, cubTemp }) if (q) cudaFree(q); }
    const char* name() const override { return "robust_nonrigid_alignment"; }
    bool capturable() const override { return csrValid; }
    unsigned long long captureKey() const override { return EnergyOps<T>::hashValue(D); }
    int grid() const { long g = ((long)(D.whi - D.wlo) + kBlock - 1) / kBlock; long cap = (long)sms * 8; if (g > cap) g = cap; if (g > kMaxPartials) g = kMaxPartials; return g < 1 ? 1 : (int)g; }
    void bind(void** q) override {   // positions: robust_nonrigid_alignment.t:5-13
        D.wf = (T) * (const float*)q[0]; D.w = (T) * (const float*)q[1];
        this->unknowns[0].user = q[2]; this->unknowns[1].user = q[3]; this->unknowns[2].user = q[4];
        D.Offset = (const T*)q[2]; D.Angle = (const T*)q[3]; D.Weights = (const T*)q[4];
        D.UrShape = (const T*)q[5]; D.Constraints = (const T*)q[6]; D.Normals = (const T*)q[7];
        const int n = *(const int*)q[8];
        if (n != nEdges || v0 != (const int*)q[9] || v1 != (const int*)q[10]) csrValid = false;
        nEdges = n; v0 = (const int*)q[9]; v1 = (const int*)q[10];
    }
    void newSolve() override { csrValid = false; }   // the caller may have rewritten the index arrays in place
    // ---- multi-GPU: 1-D decomposition of the (locality-ordered) vertex index, as energy_arap.cu: every rank binds the GLOBAL arrays, owns
    // the vertices [ownedBegin, ownedEnd) and keeps `up` / `dn` halo vertices beyond them (the graph's index bandwidth)
    int haloUp = 0, haloDn = 0;
    bool setSlab(int ownedBegin, int ownedEnd, int up, int dn) override {
        if (ownedBegin < 0 || ownedEnd > D.N || ownedBegin >= ownedEnd || ownedBegin % 4 || (ownedEnd % 4 && ownedEnd != D.N) || up % 4 || dn % 4) return false;
        if (ownedBegin - up < 0 || ownedEnd + dn > D.N || up > ownedEnd - ownedBegin || dn > ownedEnd - ownedBegin) return false;
        D.lo = ownedBegin; D.hi = ownedEnd; D.wlo = ownedBegin - up; D.whi = ownedEnd + dn; haloUp = up; haloDn = dn;
        return true;
    }
    std::vector<typename EnergyOps<T>::HaloSegment> haloSegments() const override {
        std::vector<typename EnergyOps<T>::HaloSegment> s;   // vectors are global-size on every rank: destination offsets equal source offsets
        if (haloUp > 0) { s.push_back({ 3L * D.lo, 3L * D.lo, 3L * haloUp, 0 }); s.push_back({ D.offA + 3L * D.lo, D.offA + 3L * D.lo, 3L * haloUp, 0 }); s.push_back({ D.offG + D.lo, D.offG + D.lo, (long)haloUp, 0 }); }
        if (haloDn > 0) {
            s.push_back({ 3L * (D.hi - haloDn), 3L * (D.hi - haloDn), 3L * haloDn, 1 }); s.push_back({ D.offA + 3L * (D.hi - haloDn), D.offA + 3L * (D.hi - haloDn), 3L * haloDn, 1 });
            s.push_back({ D.offG + (D.hi - haloDn), D.offG + (D.hi - haloDn), (long)haloDn, 1 });
        }
        return s;
    }
    std::vector<std::pair<long, long>> ownedScalarRanges() const override { return { { 3L * D.lo, 3L * D.hi }, { D.offA + 3L * D.lo, D.offA + 3L * D.hi }, { D.offG + D.lo, D.offG + D.hi } }; }
    std::vector<std::pair<long, long>> processScalarRanges() const override {
        if (D.wlo == 0 && D.whi == D.N) return {};
        return { { 3L * D.wlo, 3L * D.whi }, { D.offA + 3L * D.wlo, D.offA + 3L * D.whi }, { D.offG + D.wlo, D.offG + D.whi } };
    }
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
        k_rnaCsrStart<<<g, 256, 0, st>>>(nEdges, keysTmp, D.N, outStart);
        cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, v1, keysTmp, v0, inNbr, nEdges, 0, endBit, st);
        k_rnaCsrStart<<<g, 256, 0, st>>>(nEdges, keysTmp, D.N, inStart);
        D.outStart = outStart; D.outNbr = outNbr; D.inStart = inStart; D.inNbr = inNbr;
        // symmetric graphs (every mesh the reference app builds lists both directions of an edge, OptGraph.h:64-75) take the merged loop of
        // the J^T J gather: one neighbour visit serves the edge and its reverse
        int* flag = keysTmp;   // the sort keys are no longer needed
        OPT_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
        k_rnaSymmetric<<<g, 256, 0, st>>>(D.N, outStart, outNbr, inStart, flag);
        int ns = 1; OPT_CUDA_CHECK(cudaMemcpyAsync(&ns, flag, sizeof(int), cudaMemcpyDeviceToHost, st)); OPT_CUDA_CHECK(cudaStreamSynchronize(st));
        D.symmetric = (ns == 0 && !getenv("OPTB_RNA_NO_SYMMETRIC")) ? 1 : 0;
        csrValid = true;
    }
    // the pack depends on the current unknowns: refreshed wherever the reference refreshes its linearisation
    // (kernels.precompute is called before every cost evaluation and the next PCGInit1 follows it with unchanged unknowns)
    void precompute(const LaunchCtx& c) override { buildCsr(c.stream); k_rnaPack<T><<<grid(), kBlock, 0, c.stream>>>(D); }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        precompute(c);   // cheap (16 scalars per vertex) and makes pcgInit self-contained for the extension entry points
        k_rnaInit<T><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, rzSlot, this->usePreconditioner ? 1 : 0, delta, r, pre, p);
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan*) override {
        if (CtC) k_rnaApplyJTJ<T, true><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, x, out, CtC);
        else k_rnaApplyJTJ<T, false><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, dotSlot, x, out, nullptr);
    }
    void cost(const LaunchCtx& c) override { precompute(c); k_rnaCost<T, false><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, kCost, nullptr); }
    void modelCost(const LaunchCtx& c, const T* delta) override { k_rnaCost<T, true><<<grid(), kBlock, 0, c.stream>>>(D, c.ws, kModelCost, delta); }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override { k_rnaCtC<T><<<grid(), kBlock, 0, c.stream>>>(D, CtC, invRadius); }
};

template <class T> EnergyOps<T>* makeRna(const unsigned* dims) { return new RnaOps<T>(dims); }
template EnergyOps<float>* makeRna<float>(const unsigned*);
template EnergyOps<double>* makeRna<double>(const unsigned*);

}  // namespace optb
