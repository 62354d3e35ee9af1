// generic_engine.cuh -- energy-agnostic GPU back end: cached sparse Jacobian rows + deterministic gathers.
//
// For energies without a hand-tuned kernel set, the per-problem function map the reference generates
// symbolically (API/src/o.t: createjtjcentered :2029-2089, createjtjgraph :2092-2126, createjtfcentered
// :2129-2172, createjtfgraph :2228-2253, createmodelcost :2174-2225, computeCtC :2255-2316, createcost
// :2375-2385, createprecomputed :2387-2409) is realised numerically:
//   * once per linearisation (PCGInit1), the fill kernel evaluates every residual instance of the spec with
//     forward-mode dual numbers ON THE DEVICE (generic_device.cuh) and caches the residual values F and the
//     Jacobian rows (value + column) in ELL/SoA form.  ComputedArrays (kernels.precompute,
//     solverGPUGaussNewton.t:607-614) are evaluated inline by the residual code, so "precompute" is part of it.
//   * J^T F, diag(J^T J), CtC are column gathers over a transposed index (built with a stable cub radix
//     sort, so every sum has a fixed order: deterministic, no atomics -- the reference scatters graph
//     terms with atomicAdd, :687-706);  J^T J p is a row pass (y = J p) followed by a column gather.
//   * Exclude(): columns of excluded unknowns are dropped and their rows do not count in the cost, but
//     rows centred at excluded elements still feed the gathers of their included neighbours (o.t
//     residualsincludingX00 :1996-2006; SURVEY.md 8a note 2).
// The only energy-specific code is a "spec" (eval + excluded): either a C++ struct compiled into the library
// (energies_generic.cu, GenericOps<Spec,T>) or CUDA text generated from an arbitrary .t file by the front end
// and compiled with NVRTC at Opt_ProblemPlan time (frontend.cpp + dynamic_ops.cu, DynamicOps<T>).
#pragma once
#include <cub/cub.cuh>
#include <cstring>
#include <type_traits>
#include "solver.h"
#include "generic_device.cuh"

namespace optb {

// layout of the row cache shared by all kernels
struct GeLayout {
    int nGroups;
    GeGroup group[kGeMaxGroups];
    long rowBase[kGeMaxGroups + 1];   // first residual row of each group (rows of instance i: rowBase + rho*count + i)
    long entBase[kGeMaxGroups + 1];   // first cache entry of each group (entry (rho,k,i): entBase + (rho*ndeps + k)*count + i)
    long nRows, nEnt;
};

inline GeView geView(const GeLayout& L, int g) { return GeView{ L.group[g], L.rowBase[g], L.entBase[g] }; }

// ------------------------------------------------------------------ per-spec kernels (statically compiled specs)
template <class Spec, class T>
__global__ void __launch_bounds__(128) k_geFill(typename Spec::Params P, GeView V, int g, T* __restrict__ F, T* __restrict__ val, int* __restrict__ col,
                                                int* __restrict__ entRow, unsigned char* __restrict__ rowCost) { geFillBody<Spec, T>(P, V, g, F, val, col, entRow, rowCost); }
template <class Spec, class T>
__global__ void __launch_bounds__(128) k_geCost(typename Spec::Params P, GeView V, int g, double* __restrict__ partials) { geCostBody<Spec, T>(P, V, g, partials); }
template <class Spec, class T>
__global__ void __launch_bounds__(256) k_geActive(typename Spec::Params P, int a, long elements, int channels, long offset, unsigned char* __restrict__ active) {
    geActiveBody<Spec, T>(P, a, elements, channels, offset, active);
}
// sums the per-CTA partials of all groups' cost kernels in a fixed order and publishes the scalar
static __global__ void k_geFinishCost(ReduceWorkspace ws, int slot, int nPartials) {
    double t = 0;
    for (int i = threadIdx.x; i < nPartials; i += blockDim.x) t += ws.partials[(size_t)slot * kMaxPartials + i];
    __shared__ double s_w[32];
    t = warpSum(t);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double u = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) u += s_w[w];
        if (ws.comm.world <= 1) ws.scalars[slot] = u; else publishScalar(ws, slot, u);
    }
}
// keys for the transpose sort: column of each cache entry.  Unused entries and entries whose column is an excluded
// unknown get key nU (sorted to the end, never read again); colM is the masked column used by the row pass, so that
// neither pass has to look at the active flags of the columns it touches.
static __global__ void k_geKeys(long nEnt, const int* __restrict__ col, const unsigned char* __restrict__ active, int nU, int* __restrict__ keys,
                                int* __restrict__ ids, int* __restrict__ colM) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nEnt; t += (long)gridDim.x * blockDim.x) {
        int c = col[t]; if (c >= 0 && !active[c]) c = -1;
        colM[t] = c; keys[t] = c >= 0 ? c : nU; ids[t] = (int)t;
    }
}
// The masked columns and the transposed structure are built from the Exclude() flags of one linearisation; the reference
// re-evaluates the exclusion in every kernel, so a caller may change a mask image between two Opt_ProblemStep calls.
// This compares the fresh flags with the ones the structure was built from (one byte per unknown).
static __global__ void k_geActiveChanged(int nU, const unsigned char* __restrict__ now, const unsigned char* __restrict__ built, int* __restrict__ changed) {
    bool diff = false;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nU; u += gridDim.x * blockDim.x) diff = diff || (now[u] != built[u]);
    if (diff) *changed = 1;   // every writer stores the same value
}
static __global__ void k_geColStart(long nEnt, const int* __restrict__ sortedKeys, int nU, int* __restrict__ colStart) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t <= nEnt; t += (long)gridDim.x * blockDim.x) {
        const int k = t < nEnt ? sortedKeys[t] : nU, kp = t > 0 ? sortedKeys[t - 1] : -1;
        for (int u = kp + 1; u <= k && u <= nU; ++u) colStart[u] = (int)t;   // colStart[u] = first sorted position with key >= u
    }
}
// The transposed Jacobian is stored sliced-ELL (slices of 32 consecutive unknowns = one warp): slot
// sliceStart[u/32] + t*32 + u%32 holds the t-th entry of column u, so a warp walking its 32 columns reads contiguous
// 128-byte lines (a plain CSR walk costs one sector per lane per step -- measured 21 sectors/request on SFS).
static __global__ void k_geSliceWidth(int nSlices, const int* __restrict__ colStart, int* __restrict__ sliceSlots) {
    const int sl = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (sl >= nSlices) return;
    int w = colStart[sl * 32 + lane + 1] - colStart[sl * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) sliceSlots[sl] = w * 32;
}
// structure of the sliced-ELL copy: source cache entry and residual row of every slot (-1 = padding)
static __global__ void k_geSellStructure(int nU, const int* __restrict__ colStart, const int* __restrict__ sliceStart, const int* __restrict__ sortedIds,
                                         const int* __restrict__ entRow, int* __restrict__ srcS, int* __restrict__ rowS) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < nU; u += gridDim.x * blockDim.x) {
        const int s0 = colStart[u], n = colStart[u + 1] - s0, w = (sliceStart[(u >> 5) + 1] - sliceStart[u >> 5]) >> 5;
        const int base = sliceStart[u >> 5] + (u & 31);
        for (int t = 0; t < w; ++t) {
            const int id = t < n ? sortedIds[s0 + t] : -1;
            srcS[base + t * 32] = id; rowS[base + t * 32] = id >= 0 ? entRow[id] : -1;
        }
    }
}
// values of the sliced-ELL copy (one pass per linearisation)
template <class T>
__global__ void __launch_bounds__(kBlock) k_geSortValues(long nSlots, const int* __restrict__ srcS, const T* __restrict__ val, T* __restrict__ valS) {
    for (long s = (long)blockIdx.x * kBlock + threadIdx.x; s < nSlots; s += (long)gridDim.x * kBlock) { const int t = srcS[s]; valS[s] = t >= 0 ? val[t] : T(0); }
}
__device__ __forceinline__ float geGuardedInvert(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }
__device__ __forceinline__ double geGuardedInvert(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }
// walks column u of the sliced-ELL copy, eight slots in flight at a time: f(value, row)
template <class T, class F>
__device__ __forceinline__ void geWalkColumn(int u, const int* __restrict__ sliceStart, const T* __restrict__ valS, const int* __restrict__ rowS, F f) {
    const int b0 = sliceStart[u >> 5], w = (sliceStart[(u >> 5) + 1] - b0) >> 5, base = b0 + (u & 31);
    for (int t = 0; t < w; t += 8) {
        T v[8]; int r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const bool in = t + q < w; v[q] = in ? valS[base + (t + q) * 32] : T(0); r[q] = in ? rowS[base + (t + q) * 32] : -1; }
#pragma unroll
        for (int q = 0; q < 8; ++q) f(v[q], r[q]);
    }
}
// PCGInit1 (+_Graph,+_Finish), solverGPUGaussNewton.t:361-419,687-692: r = -J^T F, pre, p = pre*r, delta = 0, r.p
template <class T>
__global__ void __launch_bounds__(kBlock) k_geInit(int nU, const int* __restrict__ sliceStart, const T* __restrict__ valS,
                                                   const int* __restrict__ rowS, const T* __restrict__ F, const unsigned char* __restrict__ active,
                                                   int usePre, int graphOrder, ReduceWorkspace ws, int rzSlot, T* __restrict__ delta, T* __restrict__ r,
                                                   T* __restrict__ pre, T* __restrict__ p) {
    T acc = T(0);
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T rr = T(0), pr = T(0), pp = T(0);
        T g = T(0), d = T(0);
        geWalkColumn<T>(u, sliceStart, valS, rowS, [&](T v, int row) { const T f = row >= 0 ? F[row] : T(0); g += v * f; d += v * v; });
        if (active[u]) {
            rr = -g;
            if (graphOrder) { pr = geGuardedInvert(d); if (!usePre) pr = T(1); }        // PCGInit1_Finish order (:406-410)
            else { if (!usePre) d = T(1); pr = geGuardedInvert(d); }                    // PCGInit1 order (:379-385): the constant 1 is inverted too (=> 0.25)
            pp = pr * rr;
            acc += rr * pp;
        }
        delta[u] = T(0); r[u] = rr; pre[u] = pr; p[u] = pp;
    }
    gridReduceTo<kBlock>((double)acc, ws, rzSlot);
}
// y = J x  (row pass of PCGStep1 / computeAdelta / model cost), ALL groups in one launch, one thread per residual
// row.  The group table is staged in shared memory (indexing a by-value copy with a run-time group id would cost a
// per-thread local-memory copy).  Loads are issued in batches of eight (values, masked columns, then the gathered x)
// so that they are in flight together: a plain loop over the run-time trip count serialises one latency per entry.
template <class T>
__global__ void __launch_bounds__(kBlock) k_geJxAll(const GeLayout* __restrict__ Lp, const T* __restrict__ val, const int* __restrict__ colM,
                                                    const T* __restrict__ x, T* __restrict__ y) {
    __shared__ long s_rowBase[kGeMaxGroups + 1], s_entBase[kGeMaxGroups], s_count[kGeMaxGroups];
    __shared__ int s_ndeps[kGeMaxGroups];
    const int nG = Lp->nGroups;
    for (int g = threadIdx.x; g < nG; g += kBlock) { s_rowBase[g] = Lp->rowBase[g]; s_entBase[g] = Lp->entBase[g]; s_count[g] = Lp->group[g].count; s_ndeps[g] = Lp->group[g].ndeps; }
    if (threadIdx.x == 0) s_rowBase[nG] = Lp->rowBase[nG];
    __syncthreads();
    // 32-bit indexing throughout (bind() rejects caches with 2^31 entries or more): halves the index registers
    const unsigned nRows = (unsigned)s_rowBase[nG];
    for (unsigned row = blockIdx.x * kBlock + threadIdx.x; row < nRows; row += gridDim.x * kBlock) {
        int g = 0;
        while (g + 1 < nG && row >= (unsigned)s_rowBase[g + 1]) ++g;
        const unsigned count = (unsigned)s_count[g], local = row - (unsigned)s_rowBase[g];
        const int ndeps = s_ndeps[g];
        const unsigned rho = local / count, i = local - rho * count;
        const T* __restrict__ vp = val + ((unsigned)s_entBase[g] + rho * ndeps * count + i);
        const int* __restrict__ cp = colM + ((unsigned)s_entBase[g] + rho * ndeps * count + i);
        T s = T(0);
        for (int k0 = 0; k0 < ndeps; k0 += 8) {
            T v[8]; int c[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { const bool in = k0 + k < ndeps; v[k] = in ? vp[(unsigned)(k0 + k) * count] : T(0); c[k] = in ? cp[(unsigned)(k0 + k) * count] : -1; }
            T xv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) xv[k] = c[k] >= 0 ? x[c[k]] : T(0);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[k] * xv[k];
        }
        y[row] = s;
    }
}
// out = J^T y (+ CtC.x), dot = x.out   (column gather of PCGStep1 / computeAdelta)
template <class T>
__global__ void __launch_bounds__(kBlock) k_geJTy(int nU, const int* __restrict__ sliceStart, const T* __restrict__ valS,
                                                  const int* __restrict__ rowS, const T* __restrict__ y, const unsigned char* __restrict__ active,
                                                  const T* __restrict__ x, const T* __restrict__ CtC, ReduceWorkspace ws, int dotSlot, T* __restrict__ out) {
    T acc = T(0);
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T o = T(0);
        geWalkColumn<T>(u, sliceStart, valS, rowS, [&](T v, int row) { const T yy = row >= 0 ? y[row] : T(0); o += v * yy; });
        if (active[u]) {
            if (CtC) o += CtC[u] * x[u];
            acc += x[u] * o;
        } else o = T(0);
        out[u] = o;
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}
// PCGComputeCtC (:616-622,739-744): CtC = diag(J^T J) / radius
template <class T>
__global__ void __launch_bounds__(kBlock) k_geCtC(int nU, const int* __restrict__ sliceStart, const T* __restrict__ valS, const int* __restrict__ rowS,
                                                  const unsigned char* __restrict__ active, T invRadius, T* __restrict__ CtC) {
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T d = T(0);
        geWalkColumn<T>(u, sliceStart, valS, rowS, [&](T v, int) { d += v * v * invRadius; });
        CtC[u] = active[u] ? d : T(0);
    }
}
// computeModelCost (:666-678,746-756): 1/2 sum (F + J delta)^2 over rows that count in the cost
template <class T>
__global__ void __launch_bounds__(kBlock) k_geModelCost(long nRows, const T* __restrict__ F, const T* __restrict__ y, const unsigned char* __restrict__ rowCost,
                                                        ReduceWorkspace ws, int slot) {
    T acc = T(0);
    for (long r = (long)blockIdx.x * kBlock + threadIdx.x; r < nRows; r += (long)gridDim.x * kBlock)
        if (rowCost[r]) { const T m = F[r] + y[r]; acc += T(0.5) * m * m; }
    gridReduceTo<kBlock>((double)acc, ws, slot);
}

// ------------------------------------------------------------------ host side
// GenericOpsBase holds everything that does not depend on how the spec was compiled; the two subclasses only
// differ in how they launch the three per-spec kernels and how they bind the caller's parameter table.
template <class T>
struct GenericOpsBase : EnergyOps<T> {
    GeLayout L{};
    T *F = nullptr, *val = nullptr, *y = nullptr;
    int *col = nullptr, *colM = nullptr, *entRow = nullptr, *keys = nullptr, *ids = nullptr, *keysSorted = nullptr, *idsSorted = nullptr, *colStart = nullptr;
    // sliced-ELL transpose: sliceSlots/sliceStart per 32-unknown slice, srcS/rowS/valS per slot
    int *sliceSlots = nullptr, *sliceStart = nullptr, *srcS = nullptr, *rowS = nullptr;
    T* valS = nullptr;
    long nSlots = 0, capSlots = 0; int nSlices = 0;
    void* scanTemp = nullptr; size_t scanBytes = 0;
    GeLayout* dLayout = nullptr;
    unsigned char *rowCost = nullptr, *active = nullptr, *activeBuilt = nullptr;
    int *dChanged = nullptr, *hChanged = nullptr;
    void* cubTemp = nullptr; size_t cubBytes = 0;
    long capRows = 0, capEnt = 0;
    bool structureValid = false;
    int sms = 148;
    // ---- spec-specific hooks
    virtual void launchFill(const LaunchCtx& c, int g, int grid) = 0;
    virtual void launchCost(const LaunchCtx& c, int g, int grid, double* partials) = 0;
    virtual void launchActive(const LaunchCtx& c, int a, int grid) = 0;
    virtual void bindSpec(void** q, GeLayout& N) = 0;     // bind caller pointers; fill N.nGroups and N.group[]

    void setup() {   // after the subclass filled this->unknowns
        this->layoutUnknowns();
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        OPT_CUDA_CHECK(cudaMalloc((void**)&active, this->nUnknownsPadded)); OPT_CUDA_CHECK(cudaMemset(active, 0, this->nUnknownsPadded));
        OPT_CUDA_CHECK(cudaMalloc((void**)&activeBuilt, this->nUnknownsPadded)); OPT_CUDA_CHECK(cudaMemset(activeBuilt, 0, this->nUnknownsPadded));
        OPT_CUDA_CHECK(cudaMalloc((void**)&dChanged, sizeof(int))); OPT_CUDA_CHECK(cudaMallocHost((void**)&hChanged, sizeof(int)));
        OPT_CUDA_CHECK(cudaMalloc((void**)&colStart, sizeof(int) * (this->nUnknownsPadded + 2)));
        if (this->nUnknownsPadded % 32) { fprintf(stderr, "opt_b200: internal error, unknown vector not padded to a multiple of 32\n"); exit(2); }
        nSlices = (int)(this->nUnknownsPadded / 32);
        OPT_CUDA_CHECK(cudaMalloc((void**)&sliceSlots, sizeof(int) * (nSlices + 1))); OPT_CUDA_CHECK(cudaMemset(sliceSlots, 0, sizeof(int) * (nSlices + 1)));
        OPT_CUDA_CHECK(cudaMalloc((void**)&sliceStart, sizeof(int) * (nSlices + 1)));
        cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, sliceSlots, sliceStart, nSlices + 1);
        OPT_CUDA_CHECK(cudaMalloc(&scanTemp, scanBytes));
    }
    ~GenericOpsBase() override { for (void* q : { (void*)F, (void*)val, (void*)y, (void*)col, (void*)entRow, (void*)keys, (void*)ids, (void*)keysSorted, (void*)idsSorted, (void*)colStart, (void*)rowCost, (void*)active, (void*)rowS, (void*)valS, (void*)dLayout, cubTemp,
                                                       (void*)colM, (void*)sliceSlots, (void*)sliceStart, (void*)srcS, scanTemp, (void*)activeBuilt, (void*)dChanged }) if (q) cudaFree(q);
                                 if (hChanged) cudaFreeHost(hChanged); }
    int grid(long n, int block = kBlock) const { long g = (n + block - 1) / block; long cap = (long)sms * 8; if (g > cap) g = cap; if (g > kMaxPartials) g = kMaxPartials; return g < 1 ? 1 : (int)g; }
    void bind(void** q) override {
        GeLayout N{};
        bindSpec(q, N);
        for (int g = 0; g < N.nGroups; ++g) {
            N.rowBase[g + 1] = N.rowBase[g] + N.group[g].count * N.group[g].nres;
            N.entBase[g + 1] = N.entBase[g] + N.group[g].count * N.group[g].nres * N.group[g].ndeps;
        }
        N.nRows = N.rowBase[N.nGroups]; N.nEnt = N.entBase[N.nGroups];
        if (N.nRows != L.nRows || N.nEnt != L.nEnt) structureValid = false;
        bool layoutChanged = memcmp(&N, &L, sizeof(GeLayout)) != 0 || !dLayout;
        L = N;
        if (layoutChanged) {
            if (!dLayout) OPT_CUDA_CHECK(cudaMalloc((void**)&dLayout, sizeof(GeLayout)));
            OPT_CUDA_CHECK(cudaMemcpy(dLayout, &L, sizeof(GeLayout), cudaMemcpyHostToDevice));
        }
        if (L.nEnt >= (1L << 31)) { fprintf(stderr, "opt_b200: problem too large for the generic engine (%ld Jacobian entries)\n", L.nEnt); exit(2); }
        if (L.nRows > capRows) {
            for (void* z : { (void*)F, (void*)y, (void*)rowCost }) if (z) cudaFree(z);
            capRows = L.nRows;
            OPT_CUDA_CHECK(cudaMalloc((void**)&F, sizeof(T) * capRows)); OPT_CUDA_CHECK(cudaMalloc((void**)&y, sizeof(T) * capRows)); OPT_CUDA_CHECK(cudaMalloc((void**)&rowCost, capRows));
        }
        if (L.nEnt > capEnt) {
            for (void* z : { (void*)val, (void*)col, (void*)colM, (void*)entRow, (void*)keys, (void*)ids, (void*)keysSorted, (void*)idsSorted, cubTemp }) if (z) cudaFree(z);
            capEnt = L.nEnt;
            OPT_CUDA_CHECK(cudaMalloc((void**)&val, sizeof(T) * capEnt));
            for (int** z : { &col, &colM, &entRow, &keys, &ids, &keysSorted, &idsSorted }) OPT_CUDA_CHECK(cudaMalloc((void**)z, sizeof(int) * capEnt));
            cubBytes = 0; cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, keys, keysSorted, ids, idsSorted, (int)capEnt);
            OPT_CUDA_CHECK(cudaMalloc(&cubTemp, cubBytes));
            structureValid = false;
        }
    }
    void linearise(const LaunchCtx& c) {
        for (int a = 0; a < (int)this->unknowns.size(); ++a) launchActive(c, a, grid(this->unknowns[a].elements, 256));
        for (int g = 0; g < L.nGroups; ++g) if (L.group[g].count > 0) launchFill(c, g, grid(L.group[g].count, 128));
        // The sparsity pattern depends on the graph indices / dimensions and on the Exclude() flags; it is rebuilt at every
        // Opt_ProblemInit (newSolve), whenever the sizes change, and when a mask changed since it was built -- not per
        // nonlinear iteration.  (The stream is idle here: the previous step ended with the cost read-back.)
        if (structureValid) {
            OPT_CUDA_CHECK(cudaMemsetAsync(dChanged, 0, sizeof(int), c.stream));
            k_geActiveChanged<<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, active, activeBuilt, dChanged);
            OPT_CUDA_CHECK(cudaMemcpyAsync(hChanged, dChanged, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            OPT_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            if (*hChanged) structureValid = false;
        }
        if (!structureValid) {
            const int nU = (int)this->nUnknownsPadded;
            OPT_CUDA_CHECK(cudaMemcpyAsync(activeBuilt, active, this->nUnknownsPadded, cudaMemcpyDeviceToDevice, c.stream));
            k_geKeys<<<grid(L.nEnt), kBlock, 0, c.stream>>>(L.nEnt, col, active, nU, keys, ids, colM);
            int endBit = 1; while ((1L << endBit) <= nU && endBit < 31) ++endBit;
            cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, keys, keysSorted, ids, idsSorted, (int)L.nEnt, 0, endBit, c.stream);
            k_geColStart<<<grid(L.nEnt + 1), kBlock, 0, c.stream>>>(L.nEnt, keysSorted, nU, colStart);
            // sliced-ELL structure; its size is only known on the device, so this (once per Opt_ProblemInit) reads one int back
            k_geSliceWidth<<<(nSlices + 7) / 8, 256, 0, c.stream>>>(nSlices, colStart, sliceSlots);
            cub::DeviceScan::ExclusiveSum(scanTemp, scanBytes, sliceSlots, sliceStart, nSlices + 1, c.stream);
            int total = 0;
            OPT_CUDA_CHECK(cudaMemcpyAsync(&total, sliceStart + nSlices, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            OPT_CUDA_CHECK(cudaStreamSynchronize(c.stream));
            nSlots = total;
            if (nSlots > capSlots) {
                for (void* z : { (void*)srcS, (void*)rowS, (void*)valS }) if (z) cudaFree(z);
                capSlots = nSlots + nSlots / 8 + 1024;
                OPT_CUDA_CHECK(cudaMalloc((void**)&srcS, sizeof(int) * capSlots)); OPT_CUDA_CHECK(cudaMalloc((void**)&rowS, sizeof(int) * capSlots));
                OPT_CUDA_CHECK(cudaMalloc((void**)&valS, sizeof(T) * capSlots));
            }
            k_geSellStructure<<<grid(nU), kBlock, 0, c.stream>>>(nU, colStart, sliceStart, idsSorted, entRow, srcS, rowS);
            structureValid = true;
        }
        k_geSortValues<T><<<grid(nSlots), kBlock, 0, c.stream>>>(nSlots, srcS, val, valS);
    }
    void newSolve() override { structureValid = false; }
    unsigned long long captureKey() const override {   // buffers bind()/linearise() may reallocate + the launch geometry of the row pass
        const void* ptrs[] = { dLayout, val, colM, y, sliceStart, valS, rowS, active, F };
        unsigned long long h = EnergyOps<T>::hashBytes(ptrs, sizeof(ptrs));
        const long dims[] = { L.nRows, L.nEnt, (long)this->nUnknownsPadded, nSlots };
        return EnergyOps<T>::hashBytes(dims, sizeof(dims), h);
    }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        linearise(c);
        k_geInit<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, sliceStart, valS, rowS, F, active,
                                                                           this->usePreconditioner ? 1 : 0, this->usesGraph ? 1 : 0, c.ws, rzSlot, delta, r, pre, p);
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan*) override {
        k_geJxAll<T><<<grid(L.nRows), kBlock, 0, c.stream>>>(dLayout, val, colM, x, y);
        k_geJTy<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, sliceStart, valS, rowS, y, active, x, CtC, c.ws, dotSlot, out);
    }
    void cost(const LaunchCtx& c) override {
        // value-mode evaluation of every residual (computeCost / computeCost_Graph); per-group partials occupy
        // disjoint ranges of the partial buffer and are summed in order by one finishing CTA
        int used = 0;
        double* base = c.ws.partials + (size_t)kCost * kMaxPartials;
        for (int g = 0; g < L.nGroups; ++g) {
            if (L.group[g].count <= 0) continue;
            int gr = grid(L.group[g].count, 128); if (used + gr > kMaxPartials) gr = kMaxPartials - used;
            if (gr <= 0) { fprintf(stderr, "opt_b200: too many residual groups for the cost partial buffer\n"); exit(2); }
            launchCost(c, g, gr, base + used);
            used += gr;
        }
        k_geFinishCost<<<1, 256, 0, c.stream>>>(c.ws, kCost, used);
    }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override {
        k_geCtC<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, sliceStart, valS, rowS, active, invRadius, CtC);
    }
    void modelCost(const LaunchCtx& c, const T* delta) override {
        k_geJxAll<T><<<grid(L.nRows), kBlock, 0, c.stream>>>(dLayout, val, colM, delta, y);
        k_geModelCost<T><<<grid(L.nRows), kBlock, 0, c.stream>>>(L.nRows, F, y, rowCost, c.ws, kModelCost);
    }
};

// Spec requirements (statically compiled):  struct Params { GeArray<T> U[kGeMaxArrays]; ... };  static const char* name();
//   Spec(const unsigned* dims);  void bind(void** params, Params&);  int numArrays(); void describe(a, elements, channels);
//   int numGroups(const Params&); GeGroup group(const Params&, int g);  bool usePreconditioner; bool usesGraph;
//   template<class A> __device__ static void eval(const Params&, int g, long i, A& X, typename A::S* out);
//   __device__ static bool excluded(const Params&, int a, long e);
template <class Spec, class T>
struct GenericOps : GenericOpsBase<T> {
    Spec spec;
    typename Spec::Params P{};
    GenericOps(const unsigned* dims) : spec(dims) {
        const int na = spec.numArrays();
        this->unknowns.resize(na);
        for (int a = 0; a < na; ++a) { long el; int ch; spec.describe(a, el, ch); this->unknowns[a].elements = el; this->unknowns[a].channels = ch; }
        this->usePreconditioner = spec.usePreconditioner; this->usesGraph = spec.usesGraph;
        this->setup();
    }
    const char* name() const override { return Spec::name(); }
    void bindSpec(void** q, GeLayout& N) override {
        spec.bind(q, P);
        for (size_t a = 0; a < this->unknowns.size(); ++a) {
            this->unknowns[a].user = P.U[a].data;
            P.U[a].elements = this->unknowns[a].elements; P.U[a].channels = this->unknowns[a].channels; P.U[a].offset = this->unknowns[a].offset;
        }
        N.nGroups = spec.numGroups(P);
        for (int g = 0; g < N.nGroups; ++g) N.group[g] = spec.group(P, g);
    }
    void launchFill(const LaunchCtx& c, int g, int gr) override { k_geFill<Spec, T><<<gr, 128, 0, c.stream>>>(P, geView(this->L, g), g, this->F, this->val, this->col, this->entRow, this->rowCost); }
    void launchCost(const LaunchCtx& c, int g, int gr, double* partials) override { k_geCost<Spec, T><<<gr, 128, 0, c.stream>>>(P, geView(this->L, g), g, partials); }
    void launchActive(const LaunchCtx& c, int a, int gr) override {
        k_geActive<Spec, T><<<gr, 256, 0, c.stream>>>(P, a, this->unknowns[a].elements, this->unknowns[a].channels, this->unknowns[a].offset, this->active);
    }
};

}  // namespace optb
