// example_steps.cu -- the example harnesses' per-outer-iteration pre-steps as sm_100a kernels (include/OptB200Examples.h).
// Reference counterparts (all host loops + an upload): image_warping/src/CombinedSolver.h:181-207, arap_mesh_deformation/src/
// CombinedSolver.h:90-108, robust_nonrigid_alignment/src/CombinedSolver.h:261-346 (nanoflann kNN with MAX_K = 20, :62).
#include <algorithm>
#include <cfloat>
#include <vector>
#include <cmath>
#include <cub/cub.cuh>
#include "common.cuh"
#include "../../include/OptB200Examples.h"

namespace optb {

template <class T> __global__ void __launch_bounds__(kBlock) k_fill2(T* __restrict__ c, long n, T v) {
    for (long i = (long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long)gridDim.x * kBlock) c[i] = v;
}
template <class T> __global__ void __launch_bounds__(kBlock) k_iwMarkers(T* __restrict__ cons, const T* __restrict__ mask, int W, int H, const int* __restrict__ rows, int n, float alpha) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < n; k += gridDim.x * kBlock) {
        const int x = rows[4 * k], y = rows[4 * k + 1];
        if (x < 0 || x >= W || y < 0 || y >= H) continue;
        const long i = (long)y * W + x;
        if (mask[i] == T(0)) {   // float arithmetic like the harness (CombinedSolver.h:196-201), then stored as opt_float
            cons[2 * i] = (T)((1.0f - alpha) * (float)x + alpha * (float)rows[4 * k + 2]);
            cons[2 * i + 1] = (T)((1.0f - alpha) * (float)y + alpha * (float)rows[4 * k + 3]);
        }
    }
}
template <class T> __global__ void __launch_bounds__(kBlock) k_arapMarkers(T* __restrict__ cons, const T* __restrict__ cur, int N, const int* __restrict__ idx, const float* __restrict__ tgt, int n, float alpha) {
    for (int k = blockIdx.x * kBlock + threadIdx.x; k < n; k += gridDim.x * kBlock) {
        const int v = idx[k];
        if (v < 0 || v >= N) continue;
        for (int c = 0; c < 3; ++c) cons[3L * v + c] = (T)((1.0f - alpha) * (float)cur[3L * v + c] + alpha * tgt[3 * k + c]);
    }
}

// ---- vertex normals: unit face normals accumulated per vertex (atomics: a handful per vertex, once per outer iteration), then normalised
template <class T> __global__ void __launch_bounds__(kBlock) k_faceNormals(const T* __restrict__ pos, const int* __restrict__ faces, int nFaces, float* __restrict__ acc) {
    for (int f = blockIdx.x * kBlock + threadIdx.x; f < nFaces; f += gridDim.x * kBlock) {
        const int a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
        const float ax = (float)pos[3L * a], ay = (float)pos[3L * a + 1], az = (float)pos[3L * a + 2];
        const float ux = (float)pos[3L * b] - ax, uy = (float)pos[3L * b + 1] - ay, uz = (float)pos[3L * b + 2] - az;
        const float vx = (float)pos[3L * c] - ax, vy = (float)pos[3L * c + 1] - ay, vz = (float)pos[3L * c + 2] - az;
        float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        const float l = sqrtf(nx * nx + ny * ny + nz * nz);
        if (l > 0.f) { nx /= l; ny /= l; nz /= l; }
        for (int v : { a, b, c }) { atomicAdd(acc + 3L * v, nx); atomicAdd(acc + 3L * v + 1, ny); atomicAdd(acc + 3L * v + 2, nz); }
    }
}
__global__ void __launch_bounds__(kBlock) k_normalise3(float* __restrict__ n, int N) {
    for (int v = blockIdx.x * kBlock + threadIdx.x; v < N; v += gridDim.x * kBlock) {
        const float x = n[3L * v], y = n[3L * v + 1], z = n[3L * v + 2], l = sqrtf(x * x + y * y + z * z);
        if (l > 0.f) { n[3L * v] = x / l; n[3L * v + 1] = y / l; n[3L * v + 2] = z / l; }
    }
}

// ---- uniform grid over the target points
struct GridDesc { float ox, oy, oz, inv; int nx, ny, nz; };
__device__ __forceinline__ int cellOf(const GridDesc& g, float x, float y, float z, int* cx, int* cy, int* cz) {
    int ix = (int)floorf((x - g.ox) * g.inv), iy = (int)floorf((y - g.oy) * g.inv), iz = (int)floorf((z - g.oz) * g.inv);
    ix = min(max(ix, 0), g.nx - 1); iy = min(max(iy, 0), g.ny - 1); iz = min(max(iz, 0), g.nz - 1);
    if (cx) { *cx = ix; *cy = iy; *cz = iz; }
    return (iz * g.ny + iy) * g.nx + ix;
}
__global__ void __launch_bounds__(kBlock) k_cellKeys(GridDesc g, const float* __restrict__ p, int M, int* __restrict__ keys, int* __restrict__ ids) {
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < M; i += gridDim.x * kBlock) { keys[i] = cellOf(g, p[3L * i], p[3L * i + 1], p[3L * i + 2], nullptr, nullptr, nullptr); ids[i] = i; }
}
__global__ void __launch_bounds__(kBlock) k_cellStart(int M, const int* __restrict__ sortedKeys, int nCells, int* __restrict__ start) {
    for (int t = blockIdx.x * kBlock + threadIdx.x; t <= M; t += gridDim.x * kBlock) {
        const int k = t < M ? sortedKeys[t] : nCells, kp = t > 0 ? sortedKeys[t - 1] : -1;
        for (int u = kp + 1; u <= k && u <= nCells; ++u) start[u] = t;
    }
}
constexpr int kMaxK = 32;
template <class T>
__global__ void __launch_bounds__(128) k_rnaCorrespond(GridDesc g, const int* __restrict__ cellStart, const int* __restrict__ sortedIds, const float* __restrict__ tpos,
                                                       const float* __restrict__ tnrm, const T* __restrict__ spos, const float* __restrict__ snrm, int N, float thr, float cosThr,
                                                       int maxK, int reach, T* __restrict__ cons, T* __restrict__ cnrm, T* __restrict__ weights, float* __restrict__ prev, int* __restrict__ changed) {
    int mine = 0;
    for (int v = blockIdx.x * 128 + threadIdx.x; v < N; v += gridDim.x * 128) {
        const float px = (float)spos[3L * v], py = (float)spos[3L * v + 1], pz = (float)spos[3L * v + 2];
        float bd[kMaxK]; int bi[kMaxK]; int cnt = 0;      // the maxK nearest within thr, ascending
        int cx, cy, cz; cellOf(g, px, py, pz, &cx, &cy, &cz);
        const float thr2 = thr * thr;
        for (int dz = -reach; dz <= reach; ++dz) for (int dy = -reach; dy <= reach; ++dy) for (int dx = -reach; dx <= reach; ++dx) {
            const int qx = cx + dx, qy = cy + dy, qz = cz + dz;
            if (qx < 0 || qx >= g.nx || qy < 0 || qy >= g.ny || qz < 0 || qz >= g.nz) continue;
            const int c = (qz * g.ny + qy) * g.nx + qx;
            for (int s = cellStart[c]; s < cellStart[c + 1]; ++s) {
                const int t = sortedIds[s];
                const float ex = tpos[3L * t] - px, ey = tpos[3L * t + 1] - py, ez = tpos[3L * t + 2] - pz;
                const float d2 = ex * ex + ey * ey + ez * ez;
                if (d2 > thr2) continue;
                if (cnt == maxK && !(d2 < bd[cnt - 1] || (d2 == bd[cnt - 1] && t < bi[cnt - 1]))) continue;
                int pos = cnt < maxK ? cnt++ : maxK - 1;
                while (pos > 0 && (bd[pos - 1] > d2 || (bd[pos - 1] == d2 && bi[pos - 1] > t))) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                bd[pos] = d2; bi[pos] = t;
            }
        }
        const float sx = snrm[3L * v], sy = snrm[3L * v + 1], sz = snrm[3L * v + 2];
        int hit = -1;
        for (int k = 0; k < cnt && hit < 0; ++k) {
            const int t = bi[k];
            if (sqrtf(bd[k]) > thr) break;
            if (tnrm[3L * t] * sx + tnrm[3L * t + 1] * sy + tnrm[3L * t + 2] * sz > cosThr) hit = t;
        }
        float c3[3] = { -INFINITY, -INFINITY, -INFINITY }, n3[3] = { 0.f, 0.f, 0.f };
        if (hit >= 0) for (int c = 0; c < 3; ++c) { c3[c] = tpos[3L * hit + c]; n3[c] = tnrm[3L * hit + c]; }
        bool diff = false;
        for (int c = 0; c < 3; ++c) {
            cons[3L * v + c] = (T)c3[c];
            if (hit >= 0) cnrm[3L * v + c] = (T)n3[c];      // the harness leaves the normal of an invalid correspondence untouched (never read)
            diff = diff || !(prev[3L * v + c] == c3[c]);
            prev[3L * v + c] = c3[c];
        }
        if (diff) { weights[v] = T(1); ++mine; }
    }
    // block count -> one atomic per CTA
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(changed, s_cnt);
}

static int gridFor(long n, int block = kBlock) { long g = (n + block - 1) / block; if (g > 148 * 16) g = 148 * 16; return g < 1 ? 1 : (int)g; }

}  // namespace optb

struct OptB200_PointGrid {
    optb::GridDesc g{}; int M = 0, nCells = 0; float cell = 0;
    int *cellStart = nullptr, *sortedIds = nullptr, *dChanged = nullptr;
};

extern "C" {
using namespace optb;

int OptB200_ImageWarpingConstraints(void* dC, const void* dMask, int W, int H, const int* dRows, int n, float alpha, int dp, void* stream) {
    if (!dC || !dMask || W <= 0 || H <= 0 || (n > 0 && !dRows)) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    const long cnt = 2L * W * H;
    if (dp) { k_fill2<double><<<gridFor(cnt), kBlock, 0, st>>>((double*)dC, cnt, -1.0); if (n) k_iwMarkers<double><<<gridFor(n), kBlock, 0, st>>>((double*)dC, (const double*)dMask, W, H, dRows, n, alpha); }
    else { k_fill2<float><<<gridFor(cnt), kBlock, 0, st>>>((float*)dC, cnt, -1.0f); if (n) k_iwMarkers<float><<<gridFor(n), kBlock, 0, st>>>((float*)dC, (const float*)dMask, W, H, dRows, n, alpha); }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int OptB200_ArapConstraints(void* dC, const void* dCur, int N, const int* dIdx, const float* dT, int n, float alpha, int dp, void* stream) {
    if (!dC || !dCur || N <= 0 || (n > 0 && (!dIdx || !dT))) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    if (dp) { k_fill2<double><<<gridFor(3L * N), kBlock, 0, st>>>((double*)dC, 3L * N, -(double)INFINITY); if (n) k_arapMarkers<double><<<gridFor(n), kBlock, 0, st>>>((double*)dC, (const double*)dCur, N, dIdx, dT, n, alpha); }
    else { k_fill2<float><<<gridFor(3L * N), kBlock, 0, st>>>((float*)dC, 3L * N, -INFINITY); if (n) k_arapMarkers<float><<<gridFor(n), kBlock, 0, st>>>((float*)dC, (const float*)dCur, N, dIdx, dT, n, alpha); }
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
int OptB200_VertexNormals(const void* dPos, int isDouble, int N, const int* dFaces, int nFaces, float* dNormals, void* stream) {
    if (!dPos || !dFaces || !dNormals || N <= 0 || nFaces < 0) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(dNormals, 0, sizeof(float) * 3L * N, st);
    if (nFaces) { if (isDouble) k_faceNormals<double><<<gridFor(nFaces), kBlock, 0, st>>>((const double*)dPos, dFaces, nFaces, dNormals); else k_faceNormals<float><<<gridFor(nFaces), kBlock, 0, st>>>((const float*)dPos, dFaces, nFaces, dNormals); }
    k_normalise3<<<gridFor(N), kBlock, 0, st>>>(dNormals, N);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
OptB200_PointGrid* OptB200_PointGridCreate(const float* dT, int M, float cellSize, void* stream) {
    if (!dT || M <= 0 || !(cellSize > 0)) return nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    // bounding box on the host (once per target mesh; the harness builds its k-d tree on the host here, CombinedSolver.h:183-185)
    std::vector<float> h(3L * M);
    if (cudaMemcpyAsync(h.data(), dT, sizeof(float) * 3L * M, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return nullptr;
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int i = 0; i < M; ++i) for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], h[3L * i + c]); hi[c] = std::max(hi[c], h[3L * i + c]); }
    auto* G = new OptB200_PointGrid();
    float cell = cellSize;
    for (;;) {   // at most 2^21 cells: coarsen if the threshold is tiny relative to the extent (the query then scans more than 27 cells' worth of points, still exact)
        long nx = (long)std::floor((hi[0] - lo[0]) / cell) + 1, ny = (long)std::floor((hi[1] - lo[1]) / cell) + 1, nz = (long)std::floor((hi[2] - lo[2]) / cell) + 1;
        if (nx * ny * nz <= (1L << 21)) { G->g.nx = (int)nx; G->g.ny = (int)ny; G->g.nz = (int)nz; break; }
        cell *= 1.5f;
    }
    G->cell = cell; G->g.ox = lo[0]; G->g.oy = lo[1]; G->g.oz = lo[2]; G->g.inv = 1.0f / cell; G->M = M; G->nCells = G->g.nx * G->g.ny * G->g.nz;
    int *keys = nullptr, *ids = nullptr, *keysS = nullptr;
    cudaMalloc((void**)&keys, sizeof(int) * M); cudaMalloc((void**)&ids, sizeof(int) * M); cudaMalloc((void**)&keysS, sizeof(int) * M);
    cudaMalloc((void**)&G->sortedIds, sizeof(int) * M); cudaMalloc((void**)&G->cellStart, sizeof(int) * (G->nCells + 2)); cudaMalloc((void**)&G->dChanged, sizeof(int));
    k_cellKeys<<<gridFor(M), kBlock, 0, st>>>(G->g, dT, M, keys, ids);
    size_t tmpBytes = 0; void* tmp = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys, keysS, ids, G->sortedIds, M, 0, 22, st);
    cudaMalloc(&tmp, tmpBytes);
    cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keys, keysS, ids, G->sortedIds, M, 0, 22, st);   // stable: ids ascending inside a cell
    k_cellStart<<<gridFor(M + 1), kBlock, 0, st>>>(M, keysS, G->nCells, G->cellStart);
    cudaStreamSynchronize(st);
    cudaFree(keys); cudaFree(ids); cudaFree(keysS); cudaFree(tmp);
    if (cudaGetLastError() != cudaSuccess) { OptB200_PointGridFree(G); return nullptr; }
    return G;
}
void OptB200_PointGridFree(OptB200_PointGrid* G) { if (!G) return; cudaFree(G->cellStart); cudaFree(G->sortedIds); cudaFree(G->dChanged); delete G; }

int OptB200_RnaCorrespondences(const OptB200_PointGrid* G, const float* dTP, const float* dTN, const void* dSP, const float* dSN, int N, float thr, float cosThr, int maxK,
                               void* dC, void* dCN, void* dW, float* dPrev, int dp, void* stream) {
    if (!G || !dTP || !dTN || !dSP || !dSN || !dC || !dCN || !dW || !dPrev || N <= 0 || !(thr > 0) || maxK < 1 || maxK > kMaxK) return -3;
    cudaStream_t st = (cudaStream_t)stream;
    const int reach = (int)std::ceil(thr / G->cell);
    cudaMemsetAsync(G->dChanged, 0, sizeof(int), st);
    const int grid = gridFor(N, 128);
    if (dp) k_rnaCorrespond<double><<<grid, 128, 0, st>>>(G->g, G->cellStart, G->sortedIds, dTP, dTN, (const double*)dSP, dSN, N, thr, cosThr, maxK, reach, (double*)dC, (double*)dCN, (double*)dW, dPrev, G->dChanged);
    else k_rnaCorrespond<float><<<grid, 128, 0, st>>>(G->g, G->cellStart, G->sortedIds, dTP, dTN, (const float*)dSP, dSN, N, thr, cosThr, maxK, reach, (float*)dC, (float*)dCN, (float*)dW, dPrev, G->dChanged);
    int changed = 0;
    if (cudaMemcpyAsync(&changed, G->dChanged, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return -2;
    return changed;
}
}
