// rownorm_common.cuh — types and element helpers shared by the row-normalisation kernels (rownorm.cu: generic + wide
// TMA-staged kernels; rownorm_cluster.cu: column-split cluster kernel for layer_norm / rms_norm with affine parameters).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "axb_internal.h"
#include "reduce_common.cuh"

namespace axb {
namespace rown {

using namespace axb::red;

enum Kind { K_SOFTMAX = 0, K_LOG_SOFTMAX = 1, K_LAYER_NORM = 2, K_RMS_NORM = 3 };

struct RowParams {
    const char *in;
    char *out;
    const char *w, *b;     // may be null
    int32_t nk;            // outer dims (all dims but the axis), index 0 = innermost
    uint32_t kshape[kMaxDims];
    int64_t in_kstride[kMaxDims], out_kstride[kMaxDims];
    int64_t in_es, out_es; // byte stride along the axis
    int64_t w_es, b_es;    // byte stride of weight / bias along the axis
    int64_t rows;
    int32_t n;             // row length
    int32_t vec_ok;        // rows contiguous and suitably aligned -> vector accesses
    int32_t in_dtype, out_dtype, w_dtype, b_dtype;
    float eps;
    // wide-row (TMA-staged) kernel only: row length in 16-byte vectors, vectors per warp, smem stages, weight/bias sweep
    // order alternates between rows
    int32_t nv, vpw, stages, wb_alternate;
};

__device__ __forceinline__ void row_offsets(const RowParams &p, int64_t row, int64_t *in_off, int64_t *out_off) {
    int64_t a = 0, b = 0;
#pragma unroll 1
    for (int d = 0; d < p.nk; ++d) {
        int64_t q = row / p.kshape[d];
        int64_t c = row - q * p.kshape[d];
        a += c * p.in_kstride[d];
        b += c * p.out_kstride[d];
        row = q;
    }
    *in_off = a;
    *out_off = b;
}

template <typename T> __device__ __forceinline__ float ld_f(const char *p);
template <> __device__ __forceinline__ float ld_f<float>(const char *p) { return *reinterpret_cast<const float *>(p); }
template <> __device__ __forceinline__ float ld_f<bf16_t>(const char *p) { return bf2f(*reinterpret_cast<const uint16_t *>(p)); }
template <> __device__ __forceinline__ float ld_f<f16_t>(const char *p) { return h2f(*reinterpret_cast<const uint16_t *>(p)); }
template <typename T> __device__ __forceinline__ void st_f(char *p, float v);
template <> __device__ __forceinline__ void st_f<float>(char *p, float v) { *reinterpret_cast<float *>(p) = v; }
template <> __device__ __forceinline__ void st_f<bf16_t>(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = f2bf(v); }
template <> __device__ __forceinline__ void st_f<f16_t>(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = f2h(v); }
__device__ __forceinline__ float ld_dyn(const char *p, int dt) {
    return dt == AXB_F32 ? ld_f<float>(p) : dt == AXB_BF16 ? ld_f<bf16_t>(p) : ld_f<f16_t>(p);
}
template <typename T> __device__ __forceinline__ float round_to(float v) { return v; }
template <> __device__ __forceinline__ float round_to<bf16_t>(float v) { return bf2f(f2bf(v)); }
template <> __device__ __forceinline__ float round_to<f16_t>(float v) { return h2f(f2h(v)); }

// 4 consecutive elements
template <typename T> __device__ __forceinline__ void ld4(const char *p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const char *p, float (&v)[4]) {
    uint4 q = ldg_stream(reinterpret_cast<const uint4 *>(p));
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
}
__device__ __forceinline__ uint2 ldg64_stream(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template <> __device__ __forceinline__ void ld4<bf16_t>(const char *p, float (&v)[4]) {
    uint2 q = ldg64_stream(p);
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xFFFF0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xFFFF0000u);
}
template <> __device__ __forceinline__ void ld4<f16_t>(const char *p, float (&v)[4]) {
    uint2 q = ldg64_stream(p);
    v[0] = h2f((uint16_t)(q.x & 0xFFFFu)); v[1] = h2f((uint16_t)(q.x >> 16));
    v[2] = h2f((uint16_t)(q.y & 0xFFFFu)); v[3] = h2f((uint16_t)(q.y >> 16));
}
template <typename T> __device__ __forceinline__ void st4(char *p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(char *p, const float (&v)[4]) {
    stg_stream(reinterpret_cast<uint4 *>(p), make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                                                        __float_as_uint(v[3])));
}
__device__ __forceinline__ void stg64_stream(void *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
template <> __device__ __forceinline__ void st4<bf16_t>(char *p, const float (&v)[4]) {
    stg64_stream(p, make_uint2((uint32_t)f2bf(v[0]) | ((uint32_t)f2bf(v[1]) << 16),
                               (uint32_t)f2bf(v[2]) | ((uint32_t)f2bf(v[3]) << 16)));
}
template <> __device__ __forceinline__ void st4<f16_t>(char *p, const float (&v)[4]) {
    stg64_stream(p, make_uint2((uint32_t)f2h(v[0]) | ((uint32_t)f2h(v[1]) << 16),
                               (uint32_t)f2h(v[2]) | ((uint32_t)f2h(v[3]) << 16)));
}

struct FMax {
    __device__ static float apply(float a, float b) { return fmaxf(a, b); }
};
struct FAdd {
    __device__ static float apply(float a, float b) { return a + b; }
};
// reduce over the THREADS threads that share a row; result broadcast to all of them
template <int THREADS, class F> __device__ __forceinline__ float row_allreduce(float v, float *smem) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        v = F::apply(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    if (THREADS > 32) {
        constexpr int NW = THREADS / 32;
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        __syncthreads(); // smem reuse between consecutive reductions
        if (l == 0)
            smem[w] = v;
        __syncthreads();
        float t = smem[l % NW];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
            if (d < NW)
                t = F::apply(t, __shfl_xor_sync(0xFFFFFFFFu, t, d));
        v = t;
    }
    return v;
}


// rownorm_cluster.cu: returns true when it handled the call (status in *st)
bool try_launch_cluster(int kind, const RowParams &p, DeviceCtx *ctx, cudaStream_t s, axb_status *st);

} // namespace rown
} // namespace axb
