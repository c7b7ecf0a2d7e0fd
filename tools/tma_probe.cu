// tma_probe — which 4-D tiled tensor-map / box shapes does cp.async.bulk.tensor.4d accept? (dev tool)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int c3, uint32_t bytes, uint16_t *out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm), b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                     "l"(&m), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                     : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < bytes / 2; i += blockDim.x)
        out[i] = reinterpret_cast<uint16_t *>(sm)[i];
}
int main() {
    const uint64_t W = 56, H = 56, C = 64, N = 2;
    std::vector<uint16_t> h(W * H * C * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i & 0xFFFF);
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    struct Case { uint32_t box[4]; int swz; int c[4]; const char *name; };
    Case cases[] = {{{64, 1, 64, 1}, 1, {0, 0, 0, 0}, "box 64x1x64x1 sw128"}, {{64, 1, 64, 1}, 0, {0, 0, 0, 0}, "box 64x1x64x1 no swizzle"},
                    {{64, 2, 32, 1}, 1, {0, 0, 0, 0}, "box 64x2x32x1 sw128"}, {{64, 64, 1, 1}, 1, {0, 0, 0, 0}, "box 64x64x1x1 sw128"},
                    {{64, 1, 64, 1}, 1, {-1, -1, 0, 0}, "box 64x1x64x1 sw128 negative start"}, {{32, 1, 64, 1}, 1, {0, 0, 0, 0}, "box 32x1x64x1 sw128"},
                    {{56, 1, 64, 1}, 1, {0, 0, 0, 0}, "box 56x1x64x1 sw128"}};
    for (auto &cs : cases) {
        CUtensorMap m;
        cuuint64_t gdim[4] = {W, H, C, N}, gstr[3] = {W * 2, W * H * 2, W * H * C * 2};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, gdim, gstr, cs.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            cs.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", cs.name, (int)r); continue; }
        const uint32_t bytes = cs.box[0] * cs.box[1] * cs.box[2] * cs.box[3] * 2;
        k<<<1, 128, 32768>>>(m, cs.c[0], cs.c[1], cs.c[2], cs.c[3], bytes, o);
        cudaError_t e = cudaDeviceSynchronize();
        uint16_t first[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) cudaMemcpy(first, o, 8, cudaMemcpyDeviceToHost);
        printf("%s: %s first=%u %u %u %u\n", cs.name, cudaGetErrorString(e), first[0], first[1], first[2], first[3]);
        if (e != cudaSuccess) { printf("context dead, stop\n"); return 0; }
    }
    return 0;
}
