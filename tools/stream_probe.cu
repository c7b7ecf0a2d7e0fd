// stream_probe.cu — micro-benchmark used to choose the streaming access pattern of the elementwise family on B200
// (dev tool; results are recorded in profiles/). y = relu(x) over N fp32 with several load/store/launch variants.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
enum LD { LD_NC_NA = 0, LD_PLAIN = 1, LD_CS = 2, LD_NC = 3 };
enum ST { ST_NA = 0, ST_PLAIN = 1, ST_CS = 2 };
template <int L> __device__ __forceinline__ uint4 ld(const uint4 *p) {
    uint4 r;
    if (L == LD_NC_NA) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (L == LD_CS) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (L == LD_NC) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else r = *p;
    return r;
}
template <int S> __device__ __forceinline__ void st(uint4 *p, uint4 v) {
    if (S == ST_NA) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else if (S == ST_CS) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    else *p = v;
}
__device__ __forceinline__ uint4 relu4(uint4 v) {
    float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
    return make_uint4(__float_as_uint(a > 0 ? a : 0), __float_as_uint(b > 0 ? b : 0), __float_as_uint(c > 0 ? c : 0), __float_as_uint(d > 0 ? d : 0));
}
// each thread: U uint4 per iteration, consecutive threads consecutive uint4 (stride = total threads)
template <int L, int S, int U> __global__ void __launch_bounds__(256) k_strided(const uint4 *x, uint4 *y, size_t n4) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n4) v[u] = ld<L>(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n4) st<S>(y + i + u * stride, relu4(v[u]));
    }
}
// each thread owns 2 adjacent uint4 (32 B) like the elementwise family (VEC = 8 fp32)
template <int L, int S, int U> __global__ void __launch_bounds__(256) k_vec8(const uint4 *x, uint4 *y, size_t n8) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += U * stride) {
        uint4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n8) { v[u][0] = ld<L>(x + 2 * (i + u * stride)); v[u][1] = ld<L>(x + 2 * (i + u * stride) + 1); }
#pragma unroll
        for (int u = 0; u < U; ++u) if (i + u * stride < n8) { st<S>(y + 2 * (i + u * stride), relu4(v[u][0])); st<S>(y + 2 * (i + u * stride) + 1, relu4(v[u][1])); }
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a); for (int i = 0; i < 20; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main() {
    size_t n = (size_t)1 << 28; // 1 GiB in + 1 GiB out
    float *x, *y; cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4); cudaMemset(x, 1, n * 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double bytes = 2.0 * n * 4;
    printf("memcpy d2d: %.0f GB/s\n", bytes / timeit([&] { cudaMemcpyAsync(y, x, n * 4, cudaMemcpyDeviceToDevice); }) / 1e6);
#define RUNS(L, S, U, BPS) { int g = sms * BPS; float ms = timeit([&] { k_strided<L, S, U><<<g, 256>>>((const uint4 *)x, (uint4 *)y, n / 4); }); \
    printf("k_strided ld=%d st=%d U=%d blocks/SM=%-3d : %.0f GB/s\n", L, S, U, BPS, bytes / ms / 1e6); }
#define RUNV(L, S, U, BPS) { int g = sms * BPS; float ms = timeit([&] { k_vec8<L, S, U><<<g, 256>>>((const uint4 *)x, (uint4 *)y, n / 8); }); \
    printf("k_vec8    ld=%d st=%d U=%d blocks/SM=%-3d : %.0f GB/s\n", L, S, U, BPS, bytes / ms / 1e6); }
    RUNS(0, 0, 4, 4) RUNS(0, 0, 8, 4) RUNV(0, 0, 2, 4) RUNV(0, 0, 4, 4)
    RUNS(0, 0, 1, 8) RUNS(0, 0, 2, 8) RUNS(0, 0, 4, 8) RUNS(1, 1, 4, 8) RUNS(2, 2, 4, 8) RUNS(3, 1, 4, 8) RUNS(0, 1, 4, 8) RUNS(1, 0, 4, 8)
    RUNV(0, 0, 1, 8) RUNV(0, 0, 2, 8) RUNV(1, 1, 2, 8) RUNV(0, 0, 4, 8)
    RUNS(0, 0, 1, 16) RUNS(0, 0, 4, 16) RUNV(0, 0, 1, 16) RUNV(0, 0, 2, 16)
    RUNS(0, 0, 1, 32) RUNS(0, 0, 4, 32) RUNS(1, 1, 4, 32) RUNV(0, 0, 1, 32)
    printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    // one vector per thread, huge grid (no grid-stride)
    { size_t n4 = n / 4; unsigned g = (unsigned)((n4 + 255) / 256); float ms = timeit([&] { k_strided<0, 0, 1><<<g, 256>>>((const uint4 *)x, (uint4 *)y, n4); }); printf("one uint4 per thread, grid=%u : %.0f GB/s\n", g, bytes / ms / 1e6); }
    { size_t n4 = n / 4; unsigned g = (unsigned)((n4 / 4 + 255) / 256); float ms = timeit([&] { k_strided<0, 0, 4><<<g, 256>>>((const uint4 *)x, (uint4 *)y, n4); }); printf("4 uint4 per thread, grid=%u : %.0f GB/s\n", g, bytes / ms / 1e6); }
    return 0;
}
