// generic_engine.cuh -- energy-agnostic GPU back end: cached sparse Jacobian rows + deterministic gathers.
//
// For energies without a hand-tuned kernel set, the per-problem function map the reference generates
// symbolically (API/src/o.t: createjtjcentered :2029-2089, createjtjgraph :2092-2126, createjtfcentered
// :2129-2172, createjtfgraph :2228-2253, createmodelcost :2174-2225, computeCtC :2255-2316, createcost
// :2375-2385, createprecomputed :2387-2409) is realised numerically:
//   * once per linearisation (PCGInit1), k_geFill evaluates every residual instance of the spec with
//     forward-mode dual numbers ON THE DEVICE and caches the residual values F and the Jacobian rows
//     (value + column) in ELL/SoA form.  ComputedArrays (kernels.precompute, solverGPUGaussNewton.t
//     :607-614) are evaluated inline by the residual code, so "precompute" is part of this kernel.
//   * J^T F, diag(J^T J), CtC are column gathers over a transposed index (built with a stable cub radix
//     sort, so every sum has a fixed order: deterministic, no atomics -- the reference scatters graph
//     terms with atomicAdd, :687-706);  J^T J p is a row pass (y = J p) followed by a column gather.
//   * Exclude(): columns of excluded unknowns are dropped and their rows do not count in the cost, but
//     rows centred at excluded elements still feed the gathers of their included neighbours (o.t
//     residualsincludingX00 :1996-2006; SURVEY.md 8a note 2).
// The only energy-specific code is a small "spec" struct (energies_generic.cu) that states the residuals
// as the .t file does; a future general front end would emit exactly that piece.
#pragma once
#include <cub/cub.cuh>
#include <type_traits>
#include "solver.h"

namespace optb {

constexpr int kGeMaxDeps = 12;   // distinct unknown scalars one residual term may touch
constexpr int kGeMaxRes = 4;     // residual rows one term instance emits
constexpr int kGeMaxGroups = 6;
constexpr int kGeMaxArrays = 4;

// ------------------------------------------------------------------ device forward-mode dual numbers
template <class T>
struct GDual {
    T v; T d[kGeMaxDeps];
    __device__ GDual() : v(0) { for (int i = 0; i < kGeMaxDeps; ++i) d[i] = T(0); }
    __device__ GDual(T c) : v(c) { for (int i = 0; i < kGeMaxDeps; ++i) d[i] = T(0); }
};
#define GE_BIN(op, VEXP, DEXP)                                                                       \
    template <class T> __device__ __forceinline__ GDual<T> operator op(const GDual<T>& a, const GDual<T>& b) { \
        GDual<T> r; r.v = VEXP;                                                                      \
        _Pragma("unroll") for (int i = 0; i < kGeMaxDeps; ++i) r.d[i] = DEXP;                        \
        return r; }
GE_BIN(+, a.v + b.v, a.d[i] + b.d[i])
GE_BIN(-, a.v - b.v, a.d[i] - b.d[i])
GE_BIN(*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
GE_BIN(/, a.v / b.v, (a.d[i] * b.v - a.v * b.d[i]) / (b.v * b.v))
#undef GE_BIN
template <class T> __device__ __forceinline__ GDual<T> operator+(const GDual<T>& a, T b) { GDual<T> r = a; r.v += b; return r; }
template <class T> __device__ __forceinline__ GDual<T> operator+(T a, const GDual<T>& b) { GDual<T> r = b; r.v += a; return r; }
template <class T> __device__ __forceinline__ GDual<T> operator-(const GDual<T>& a, T b) { GDual<T> r = a; r.v -= b; return r; }
template <class T> __device__ __forceinline__ GDual<T> operator-(T a, const GDual<T>& b) { return GDual<T>(a) - b; }
template <class T> __device__ __forceinline__ GDual<T> operator*(const GDual<T>& a, T b) { GDual<T> r; r.v = a.v * b; _Pragma("unroll") for (int i = 0; i < kGeMaxDeps; ++i) r.d[i] = a.d[i] * b; return r; }
template <class T> __device__ __forceinline__ GDual<T> operator*(T a, const GDual<T>& b) { return b * a; }
template <class T> __device__ __forceinline__ GDual<T> operator/(const GDual<T>& a, T b) { return a * (T(1) / b); }
template <class T> __device__ __forceinline__ GDual<T> operator/(T a, const GDual<T>& b) { return GDual<T>(a) / b; }
template <class T> __device__ __forceinline__ GDual<T> operator-(const GDual<T>& a) { return a * T(-1); }
__device__ __forceinline__ float geSin(float x) { return sinf(x); }
__device__ __forceinline__ double geSin(double x) { return sin(x); }
__device__ __forceinline__ float geCos(float x) { return cosf(x); }
__device__ __forceinline__ double geCos(double x) { return cos(x); }
__device__ __forceinline__ float geSqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double geSqrt(double x) { return sqrt(x); }
template <class T> __device__ __forceinline__ GDual<T> geSin(const GDual<T>& a) { GDual<T> r; r.v = geSin(a.v); const T c = geCos(a.v); _Pragma("unroll") for (int i = 0; i < kGeMaxDeps; ++i) r.d[i] = c * a.d[i]; return r; }
template <class T> __device__ __forceinline__ GDual<T> geCos(const GDual<T>& a) { GDual<T> r; r.v = geCos(a.v); const T s = -geSin(a.v); _Pragma("unroll") for (int i = 0; i < kGeMaxDeps; ++i) r.d[i] = s * a.d[i]; return r; }
template <class T> __device__ __forceinline__ GDual<T> geSqrt(const GDual<T>& a) { GDual<T> r; r.v = geSqrt(a.v); const T s = T(0.5) / r.v; _Pragma("unroll") for (int i = 0; i < kGeMaxDeps; ++i) r.d[i] = s * a.d[i]; return r; }
template <class T> __device__ __forceinline__ T geVal(const GDual<T>& a) { return a.v; }
__device__ __forceinline__ float geVal(float a) { return a; }
__device__ __forceinline__ double geVal(double a) { return a; }

// ------------------------------------------------------------------ spec-facing types
template <class T> struct GeArray { T* data; long elements; int channels; long offset; };   // one Unknown(...) array (flat offset = plan layout)
struct GeGroup { long count; int nres; int ndeps; int centered; int centerArray; };          // one Energy(...) statement over its domain

template <class T> struct GeValAcc {   // plain evaluation
    using S = T;
    const GeArray<T>* U;
    __device__ __forceinline__ S operator()(int a, long e, int c) const { return U[a].data[e * U[a].channels + c]; }
};
template <class T> struct GeDualAcc {  // assigns each distinct unknown scalar touched a derivative slot
    using S = GDual<T>;
    const GeArray<T>* U;
    int col[kGeMaxDeps]; int n;
    __device__ __forceinline__ S operator()(int a, long e, int c) {
        const int g = (int)(U[a].offset + e * U[a].channels + c);
        int s = -1;
#pragma unroll
        for (int i = 0; i < kGeMaxDeps; ++i) if (i < n && col[i] == g) s = i;
        if (s < 0) {
            s = n < kGeMaxDeps ? n : kGeMaxDeps - 1;
#pragma unroll
            for (int i = 0; i < kGeMaxDeps; ++i) if (i == n) col[i] = g;   // select, not a dynamic store: keeps col[] in registers
            if (n < kGeMaxDeps) ++n;
        }
        S r(U[a].data[e * U[a].channels + c]);
#pragma unroll
        for (int i = 0; i < kGeMaxDeps; ++i) if (i == s) r.d[i] = T(1);
        return r;
    }
};

// layout of the row cache shared by all kernels
struct GeLayout {
    int nGroups;
    GeGroup group[kGeMaxGroups];
    long rowBase[kGeMaxGroups + 1];   // first residual row of each group (rows of instance i: rowBase + rho*count + i)
    long entBase[kGeMaxGroups + 1];   // first cache entry of each group (entry (rho,k,i): entBase + (rho*ndeps + k)*count + i)
    long nRows, nEnt;
};

// one group's slice of the layout, passed by value (indexing the by-value GeLayout with a run-time group id
// would force a per-thread local-memory copy of the whole struct)
struct GeView { GeGroup G; long rowBase, entBase; };
inline GeView geView(const GeLayout& L, int g) { return GeView{ L.group[g], L.rowBase[g], L.entBase[g] }; }

// ------------------------------------------------------------------ kernels
template <class Spec, class T>
__global__ void __launch_bounds__(128) k_geFill(typename Spec::Params P, GeView V, int g, T* __restrict__ F, T* __restrict__ val, int* __restrict__ col,
                                                int* __restrict__ entRow, unsigned char* __restrict__ rowCost) {
    const GeGroup G = V.G;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < G.count; i += (long)gridDim.x * blockDim.x) {
        GeDualAcc<T> X; X.U = P.U; X.n = 0;
#pragma unroll
        for (int k = 0; k < kGeMaxDeps; ++k) X.col[k] = -1;
        GDual<T> out[kGeMaxRes];
        Spec::eval(P, g, i, X, out);
        const unsigned char rc = (G.centered && Spec::excluded(P, G.centerArray, i)) ? 0 : 1;
        for (int rho = 0; rho < G.nres; ++rho) {
            const long row = V.rowBase + (long)rho * G.count + i;
            F[row] = out[rho].v; rowCost[row] = rc;
            for (int k = 0; k < G.ndeps; ++k) {
                const long e = V.entBase + ((long)rho * G.ndeps + k) * G.count + i;
                T v = T(0);
#pragma unroll
                for (int q = 0; q < kGeMaxDeps; ++q) if (q == k) v = out[rho].d[q];
                int c = -1;
#pragma unroll
                for (int q = 0; q < kGeMaxDeps; ++q) if (q == k && q < X.n) c = X.col[q];
                val[e] = c >= 0 ? v : T(0); col[e] = c; entRow[e] = (int)row;
            }
        }
    }
}
template <class Spec, class T>
__global__ void __launch_bounds__(128) k_geCost(typename Spec::Params P, GeView V, int g, ReduceWorkspace ws, int slot) {
    const GeGroup G = V.G;
    T acc = T(0);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < G.count; i += (long)gridDim.x * blockDim.x) {
        if (G.centered && Spec::excluded(P, G.centerArray, i)) continue;
        GeValAcc<T> X; X.U = P.U;
        T out[kGeMaxRes];
        Spec::eval(P, g, i, X, out);
        for (int rho = 0; rho < G.nres; ++rho) acc += T(0.5) * out[rho] * out[rho];
    }
    __shared__ double s_w[4];
    double v = warpSum((double)acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w]; ws.partials[(size_t)slot * kMaxPartials + blockIdx.x] = t; }
}
// sums the per-CTA partials of all groups' cost kernels in a fixed order and publishes the scalar
__global__ void k_geFinishCost(ReduceWorkspace ws, int slot, int nPartials) {
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
template <class Spec, class T>
__global__ void __launch_bounds__(256) k_geActive(typename Spec::Params P, int nArrays, unsigned char* __restrict__ active) {
    for (int a = 0; a < nArrays; ++a) {
        const GeArray<T> A = P.U[a];
        for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < A.elements; e += (long)gridDim.x * blockDim.x) {
            const unsigned char act = Spec::excluded(P, a, e) ? 0 : 1;
            for (int c = 0; c < A.channels; ++c) active[A.offset + e * A.channels + c] = act;
        }
    }
}
// keys for the transpose sort: column of each cache entry (unused entries -> nU, sorted to the end)
__global__ void k_geKeys(long nEnt, const int* __restrict__ col, int nU, int* __restrict__ keys, int* __restrict__ ids) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nEnt; t += (long)gridDim.x * blockDim.x) { keys[t] = col[t] >= 0 ? col[t] : nU; ids[t] = (int)t; }
}
__global__ void k_geColStart(long nEnt, const int* __restrict__ sortedKeys, int nU, int* __restrict__ colStart) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t <= nEnt; t += (long)gridDim.x * blockDim.x) {
        const int k = t < nEnt ? sortedKeys[t] : nU, kp = t > 0 ? sortedKeys[t - 1] : -1;
        for (int u = kp + 1; u <= k && u <= nU; ++u) colStart[u] = (int)t;   // colStart[u] = first sorted position with key >= u
    }
}
__device__ __forceinline__ float geGuardedInvert(float v) { float s = 1.0f + sqrtf(v); return 1.0f / (s * s); }
__device__ __forceinline__ double geGuardedInvert(double v) { double s = 1.0 + sqrt(v); return 1.0 / (s * s); }
// PCGInit1 (+_Graph,+_Finish), solverGPUGaussNewton.t:361-419,687-692: r = -J^T F, pre, p = pre*r, delta = 0, r.p
template <class T>
__global__ void __launch_bounds__(kBlock) k_geInit(int nU, const int* __restrict__ colStart, const int* __restrict__ sortedIds, const T* __restrict__ val,
                                                   const int* __restrict__ entRow, const T* __restrict__ F, const unsigned char* __restrict__ active,
                                                   int usePre, int graphOrder, ReduceWorkspace ws, int rzSlot, T* __restrict__ delta, T* __restrict__ r,
                                                   T* __restrict__ pre, T* __restrict__ p) {
    T acc = T(0);
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T rr = T(0), pr = T(0), pp = T(0);
        if (active[u]) {
            T g = T(0), d = T(0);
            for (int s = colStart[u]; s < colStart[u + 1]; ++s) { const int t = sortedIds[s]; const T v = val[t]; g += v * F[entRow[t]]; d += v * v; }
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
// y = J x  (row pass of PCGStep1 / computeAdelta / model cost)
template <class T>
__global__ void __launch_bounds__(kBlock) k_geJx(GeView V, const T* __restrict__ val, const int* __restrict__ col, const unsigned char* __restrict__ active,
                                                 const T* __restrict__ x, T* __restrict__ y) {
    const GeGroup G = V.G;
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < G.count; i += (long)gridDim.x * kBlock) {
        for (int rho = 0; rho < G.nres; ++rho) {
            T s = T(0);
            for (int k = 0; k < G.ndeps; ++k) {
                const long e = V.entBase + ((long)rho * G.ndeps + k) * G.count + i;
                const int c = col[e];
                if (c >= 0 && active[c]) s += val[e] * x[c];
            }
            y[V.rowBase + (long)rho * G.count + i] = s;
        }
    }
}
// out = J^T y (+ CtC.x), dot = x.out   (column gather of PCGStep1 / computeAdelta)
template <class T>
__global__ void __launch_bounds__(kBlock) k_geJTy(int nU, const int* __restrict__ colStart, const int* __restrict__ sortedIds, const T* __restrict__ val,
                                                  const int* __restrict__ entRow, const T* __restrict__ y, const unsigned char* __restrict__ active,
                                                  const T* __restrict__ x, const T* __restrict__ CtC, ReduceWorkspace ws, int dotSlot, T* __restrict__ out) {
    T acc = T(0);
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T o = T(0);
        if (active[u]) {
            for (int s = colStart[u]; s < colStart[u + 1]; ++s) { const int t = sortedIds[s]; o += val[t] * y[entRow[t]]; }
            if (CtC) o += CtC[u] * x[u];
            acc += x[u] * o;
        }
        out[u] = o;
    }
    if (dotSlot >= 0) gridReduceTo<kBlock>((double)acc, ws, dotSlot);
}
// PCGComputeCtC (:616-622,739-744): CtC = diag(J^T J) / radius
template <class T>
__global__ void __launch_bounds__(kBlock) k_geCtC(int nU, const int* __restrict__ colStart, const int* __restrict__ sortedIds, const T* __restrict__ val,
                                                  const unsigned char* __restrict__ active, T invRadius, T* __restrict__ CtC) {
    for (int u = blockIdx.x * kBlock + threadIdx.x; u < nU; u += gridDim.x * kBlock) {
        T d = T(0);
        if (active[u]) for (int s = colStart[u]; s < colStart[u + 1]; ++s) { const T v = val[sortedIds[s]]; d += v * v * invRadius; }
        CtC[u] = d;
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

// ------------------------------------------------------------------ host side: EnergyOps over a Spec
// Spec requirements:  struct Params { GeArray<T> U[kGeMaxArrays]; ... };  static const char* name();
//   Spec(const unsigned* dims);  void bind(void** params, Params&);  int numArrays(); arrays described by describe(a, elements, channels);
//   int numGroups(const Params&); GeGroup group(const Params&, int g);  bool usePreconditioner; bool usesGraph;
//   template<class A> __device__ static void eval(const Params&, int g, long i, A& X, typename A::S* out);
//   __device__ static bool excluded(const Params&, int a, long e);
template <class Spec, class T>
struct GenericOps : EnergyOps<T> {
    Spec spec;
    typename Spec::Params P{};
    GeLayout L{};
    T *F = nullptr, *val = nullptr, *y = nullptr;
    int *col = nullptr, *entRow = nullptr, *keys = nullptr, *ids = nullptr, *keysSorted = nullptr, *idsSorted = nullptr, *colStart = nullptr;
    unsigned char *rowCost = nullptr, *active = nullptr;
    void* cubTemp = nullptr; size_t cubBytes = 0;
    long capRows = 0, capEnt = 0;
    bool structureValid = false;
    int sms = 148;
    GenericOps(const unsigned* dims) : spec(dims) {
        const int na = spec.numArrays();
        this->unknowns.resize(na);
        for (int a = 0; a < na; ++a) { long el; int ch; spec.describe(a, el, ch); this->unknowns[a].elements = el; this->unknowns[a].channels = ch; }
        this->layoutUnknowns();
        this->usePreconditioner = spec.usePreconditioner; this->usesGraph = spec.usesGraph;
        int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        OPT_CUDA_CHECK(cudaMalloc((void**)&active, this->nUnknownsPadded)); OPT_CUDA_CHECK(cudaMemset(active, 0, this->nUnknownsPadded));
        OPT_CUDA_CHECK(cudaMalloc((void**)&colStart, sizeof(int) * (this->nUnknownsPadded + 2)));
    }
    ~GenericOps() override { for (void* q : { (void*)F, (void*)val, (void*)y, (void*)col, (void*)entRow, (void*)keys, (void*)ids, (void*)keysSorted, (void*)idsSorted, (void*)colStart, (void*)rowCost, (void*)active, cubTemp }) if (q) cudaFree(q); }
    const char* name() const override { return Spec::name(); }
    int grid(long n, int block = kBlock) const { long g = (n + block - 1) / block; long cap = (long)sms * 8; if (g > cap) g = cap; if (g > kMaxPartials) g = kMaxPartials; return g < 1 ? 1 : (int)g; }
    void bind(void** q) override {
        spec.bind(q, P);
        for (size_t a = 0; a < this->unknowns.size(); ++a) {
            this->unknowns[a].user = P.U[a].data;
            P.U[a].elements = this->unknowns[a].elements; P.U[a].channels = this->unknowns[a].channels; P.U[a].offset = this->unknowns[a].offset;
        }
        GeLayout N{}; N.nGroups = spec.numGroups(P);
        for (int g = 0; g < N.nGroups; ++g) {
            N.group[g] = spec.group(P, g);
            N.rowBase[g + 1] = N.rowBase[g] + N.group[g].count * N.group[g].nres;
            N.entBase[g + 1] = N.entBase[g] + N.group[g].count * N.group[g].nres * N.group[g].ndeps;
        }
        N.nRows = N.rowBase[N.nGroups]; N.nEnt = N.entBase[N.nGroups];
        if (N.nRows != L.nRows || N.nEnt != L.nEnt) structureValid = false;
        L = N;
        if (L.nEnt >= (1L << 31)) { fprintf(stderr, "opt_b200: problem too large for the generic engine (%ld Jacobian entries)\n", L.nEnt); exit(2); }
        if (L.nRows > capRows) {
            for (void* z : { (void*)F, (void*)y, (void*)rowCost }) if (z) cudaFree(z);
            capRows = L.nRows;
            OPT_CUDA_CHECK(cudaMalloc((void**)&F, sizeof(T) * capRows)); OPT_CUDA_CHECK(cudaMalloc((void**)&y, sizeof(T) * capRows)); OPT_CUDA_CHECK(cudaMalloc((void**)&rowCost, capRows));
        }
        if (L.nEnt > capEnt) {
            for (void* z : { (void*)val, (void*)col, (void*)entRow, (void*)keys, (void*)ids, (void*)keysSorted, (void*)idsSorted, cubTemp }) if (z) cudaFree(z);
            capEnt = L.nEnt;
            OPT_CUDA_CHECK(cudaMalloc((void**)&val, sizeof(T) * capEnt));
            for (int** z : { &col, &entRow, &keys, &ids, &keysSorted, &idsSorted }) OPT_CUDA_CHECK(cudaMalloc((void**)z, sizeof(int) * capEnt));
            cubBytes = 0; cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, keys, keysSorted, ids, idsSorted, (int)capEnt);
            OPT_CUDA_CHECK(cudaMalloc(&cubTemp, cubBytes));
            structureValid = false;
        }
    }
    void linearise(const LaunchCtx& c) {
        const int na = (int)this->unknowns.size();
        k_geActive<Spec, T><<<grid(this->nUnknownsPadded), 256, 0, c.stream>>>(P, na, active);
        for (int g = 0; g < L.nGroups; ++g)
            k_geFill<Spec, T><<<grid(L.group[g].count, 128), 128, 0, c.stream>>>(P, geView(L, g), g, F, val, col, entRow, rowCost);
        // The sparsity pattern depends on the graph indices / dimensions only; it is rebuilt at every Opt_ProblemInit
        // (invalidate()) and whenever the sizes change, not per nonlinear iteration.
        if (!structureValid) {
            const int nU = (int)this->nUnknownsPadded;
            k_geKeys<<<grid(L.nEnt), kBlock, 0, c.stream>>>(L.nEnt, col, nU, keys, ids);
            int endBit = 1; while ((1L << endBit) <= nU && endBit < 31) ++endBit;
            cub::DeviceRadixSort::SortPairs(cubTemp, cubBytes, keys, keysSorted, ids, idsSorted, (int)L.nEnt, 0, endBit, c.stream);
            k_geColStart<<<grid(L.nEnt + 1), kBlock, 0, c.stream>>>(L.nEnt, keysSorted, nU, colStart);
            structureValid = true;
        }
    }
    void invalidate() { structureValid = false; }
    void pcgInit(const LaunchCtx& c, T* delta, T* r, T* pre, T* p, int rzSlot) override {
        linearise(c);
        k_geInit<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, colStart, idsSorted, val, entRow, F, active,
                                                                           this->usePreconditioner ? 1 : 0, this->usesGraph ? 1 : 0, c.ws, rzSlot, delta, r, pre, p);
    }
    void applyJTJ(const LaunchCtx& c, const T* x, T* out, const T* CtC, int dotSlot, const HaloPlan*) override {
        for (int g = 0; g < L.nGroups; ++g) k_geJx<T><<<grid(L.group[g].count), kBlock, 0, c.stream>>>(geView(L, g), val, col, active, x, y);
        k_geJTy<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, colStart, idsSorted, val, entRow, y, active, x, CtC, c.ws, dotSlot, out);
    }
    void cost(const LaunchCtx& c) override {
        // value-mode evaluation of every residual (computeCost / computeCost_Graph); per-group partials occupy
        // disjoint ranges of the partial buffer and are summed in order by one finishing CTA
        int used = 0;
        for (int g = 0; g < L.nGroups; ++g) {
            int gr = grid(L.group[g].count, 128); if (used + gr > kMaxPartials) gr = kMaxPartials - used;
            ReduceWorkspace w = c.ws; w.partials = c.ws.partials + used;   // shifted view: kernel writes partials[slot*kMax + blockIdx]
            k_geCost<Spec, T><<<gr, 128, 0, c.stream>>>(P, geView(L, g), g, w, kCost);
            used += gr;
        }
        k_geFinishCost<<<1, 256, 0, c.stream>>>(c.ws, kCost, used);
    }
    void computeCtC(const LaunchCtx& c, T* CtC, T invRadius) override {
        k_geCtC<T><<<grid(this->nUnknownsPadded), kBlock, 0, c.stream>>>((int)this->nUnknownsPadded, colStart, idsSorted, val, active, invRadius, CtC);
    }
    void modelCost(const LaunchCtx& c, const T* delta) override {
        for (int g = 0; g < L.nGroups; ++g) k_geJx<T><<<grid(L.group[g].count), kBlock, 0, c.stream>>>(geView(L, g), val, col, active, delta, y);
        k_geModelCost<T><<<grid(L.nRows), kBlock, 0, c.stream>>>(L.nRows, F, y, rowCost, c.ws, kModelCost);
    }
    void newSolve() override { invalidate(); }
};

}  // namespace optb
