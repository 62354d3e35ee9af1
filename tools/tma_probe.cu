// standalone probe: which tensor-map shapes does UTMALDG accept?  (debugging aid, not product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m, int bytes, int x, int y, float* out, int n) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(sa(smem)), "l"(&m), "r"(sa(&bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(sa(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    void* fn; cudaDriverEntryPointQueryResult qr; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    EncodeFn enc = (EncodeFn)fn;
    struct Case { int W, H, bw, bh, x, y; } cases[] = { {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6])} };
    for (auto c : cases) {
        float* d; cudaMalloc(&d, (size_t)c.W * c.H * 4);
        std::vector<float> h((size_t)c.W * c.H); for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
        cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)c.W, (cuuint64_t)c.H}, st[1] = {(cuuint64_t)c.W * 4}; cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh}, es[2] = {1, 1};
        CUresult rc = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int n = c.bw * c.bh; float* o; cudaMalloc(&o, n * 4);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        k<<<1, 128, 65536>>>(m, n * 4, c.x, c.y, o, n);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> r(n); cudaMemcpy(r.data(), o, n * 4, cudaMemcpyDeviceToHost);
        printf("W=%d H=%d box=%dx%d at(%d,%d): encode=%d run=%s  first row: %g %g %g %g | row1: %g %g %g\n", c.W, c.H, c.bw, c.bh, c.x, c.y, (int)rc, cudaGetErrorString(e), r[0], r[1], r[2], r[3], r[c.bw], r[c.bw + 1], r[c.bw + 2]);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
