// pool_sort.cu — pooling, sort / argsort / topk and masked_select: the last operators of SURVEY.md §8f row 4 ("remaining GPU
// op surface") that still round-tripped through the host. The reference has no device branch for any of them: pooling and
// sorting copy the tensor to the CPU, run scalar loops and upload the result (src/tensor/operations.cpp:2367-2880 pooling,
// :4053-4257 sort / argsort / topk); masked_select is a registry operation with a CPU implementation only
// (src/backends/cpu/cpu_operations.cpp:2719-2789).
//
//   pooling        one thread per OUTPUT element, consecutive threads along the innermost output dim; the window is walked
//                  in the CPU loops' order (kd, kh, kw) with their accumulation type (fp64 max / fp64 sum, sum / count in
//                  fp64, one rounding through float for the half types) — results are bit-identical for every float dtype.
//   sort family    bitonic network over order-preserving unsigned keys paired with the element's position, so the result is
//                  the STABLE order (ties by ascending position) for both directions; the reference uses std::sort /
//                  std::partial_sort, whose order among equal keys is unspecified, so values always agree and indices agree
//                  wherever the keys are distinct. A slice that fits one CTA's shared memory is loaded, sorted and stored
//                  by one kernel; longer slices run the classic global / shared two-level schedule over a scratch buffer.
//                  topk = the first k of the sorted slice (std::partial_sort leaves them sorted, so `sorted` is immaterial).
//   masked_select  count per 2048-element chunk -> one-CTA exclusive scan -> ordered compaction; the element count is the one
//                  8-byte device->host read of the call (the result's shape depends on it). The mask is "first byte of the
//                  element != 0", whatever its dtype, as in the CPU loop (:2738-2741).
#include <algorithm>
#include <cstring>

#include "axb_internal.h"
#include "reduce_common.cuh"

namespace axb {
namespace {

// ---- element access shared by the pooling kernels (same conversions as spatial.cu) -------------------------------------
template <int DT> struct PElem;
template <> struct PElem<AXB_F32> {
    static __device__ __forceinline__ double ld(const char *p) { return (double)*reinterpret_cast<const float *>(p); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<float *>(p) = (float)v; }
};
template <> struct PElem<AXB_F64> {
    static __device__ __forceinline__ double ld(const char *p) { return *reinterpret_cast<const double *>(p); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<double *>(p) = v; }
};
template <> struct PElem<AXB_BF16> { // double -> float -> bf16 (bfloat16.hpp:29)
    static __device__ __forceinline__ double ld(const char *p) { return (double)red::bf2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<uint16_t *>(p) = red::f2bf((float)v); }
};
template <> struct PElem<AXB_F16> { // float16.hpp:16
    static __device__ __forceinline__ double ld(const char *p) { return (double)red::h2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<uint16_t *>(p) = red::f2h((float)v); }
};

struct PoolParams {
    const char *in;
    char *out;
    int64_t in_s[5]; // byte strides of (N, C, D, H, W)
    int64_t C, D, H, W, OD, OH, OW;
    int32_t k[3], s[3], p[3];
    int32_t include_pad, esize;
    int64_t total;
};

template <int DT, bool IS_AVG> __global__ void __launch_bounds__(256) pool_kernel(const __grid_constant__ PoolParams q) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ow = i % q.OW, t1 = i / q.OW, oh = t1 % q.OH, t2 = t1 / q.OH, od = t2 % q.OD, t3 = t2 / q.OD, c = t3 % q.C,
                      n = t3 / q.C;
        const char *base = q.in + n * q.in_s[0] + c * q.in_s[1];
        const int64_t d0 = od * q.s[0] - q.p[0], h0 = oh * q.s[1] - q.p[1], w0 = ow * q.s[2] - q.p[2];
        double acc = IS_AVG ? 0.0 : -__longlong_as_double(0x7FF0000000000000ll);
        int count = 0;
        for (int kd = 0; kd < q.k[0]; ++kd) {
            const int64_t d = d0 + kd;
            const bool din = d >= 0 && d < q.D;
            for (int kh = 0; kh < q.k[1]; ++kh) {
                const int64_t h = h0 + kh;
                const bool hin = din && h >= 0 && h < q.H;
                for (int kw = 0; kw < q.k[2]; ++kw) {
                    const int64_t w = w0 + kw;
                    if (hin && w >= 0 && w < q.W) {
                        const double v = PElem<DT>::ld(base + d * q.in_s[2] + h * q.in_s[3] + w * q.in_s[4]);
                        if (IS_AVG) {
                            acc += v;
                            ++count;
                        } else if (v > acc) { // NaN never wins (operations.cpp:2407-2408)
                            acc = v;
                        }
                    } else if (IS_AVG && q.include_pad) {
                        ++count;
                    }
                }
            }
        }
        if (IS_AVG)
            acc = count > 0 ? acc / (double)count : 0.0;
        PElem<DT>::st(q.out + i * q.esize, acc);
    }
}

template <int DT> axb_status launch_pool(const PoolParams &q, bool avg, cudaStream_t s, int sms) {
    const int threads = 256;
    const int64_t want = (q.total + threads - 1) / threads;
    const int blocks = (int)std::min<int64_t>(want, (int64_t)sms * 16);
    if (avg)
        pool_kernel<DT, true><<<blocks, threads, 0, s>>>(q);
    else
        pool_kernel<DT, false><<<blocks, threads, 0, s>>>(q);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

// ---- sort ---------------------------------------------------------------------------------------------------------------
// order-preserving unsigned image of a key: a < b  <=>  image(a) < image(b) (floats: -0 is keyed as +0, NaNs sit at the
// two ends by sign); descending = the complement
__device__ __forceinline__ uint32_t image32f(uint32_t u) { return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ uint64_t image64f(uint64_t u) { return u ^ ((u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull); }

struct E8 { // 32-bit key and 32-bit position in one word
    unsigned long long v;
};
struct E16 {
    unsigned long long key, idx;
};
__device__ __forceinline__ bool lt(const E8 &a, const E8 &b) { return a.v < b.v; }
__device__ __forceinline__ bool lt(const E16 &a, const E16 &b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); }

template <int DT> struct SortKey;
template <> struct SortKey<AXB_F32> {
    using E = E8;
    using T = uint32_t;
    static __device__ __forceinline__ E make(T raw, uint64_t pos, bool desc) {
        uint32_t k = image32f(raw == 0x80000000u ? 0u : raw); // -0 == +0 under `<` (ties keep their positions)
        if (desc)
            k = ~k;
        return E{((unsigned long long)k << 32) | (unsigned long long)(uint32_t)pos};
    }
    static __device__ __forceinline__ int64_t index(const E &e) { return (int64_t)(uint32_t)e.v; }
    static __device__ __forceinline__ E sentinel() { return E{~0ull}; }
};
template <> struct SortKey<AXB_I32> {
    using E = E8;
    using T = uint32_t;
    static __device__ __forceinline__ E make(T raw, uint64_t pos, bool desc) {
        uint32_t k = raw ^ 0x80000000u;
        if (desc)
            k = ~k;
        return E{((unsigned long long)k << 32) | (unsigned long long)(uint32_t)pos};
    }
    static __device__ __forceinline__ int64_t index(const E &e) { return (int64_t)(uint32_t)e.v; }
    static __device__ __forceinline__ E sentinel() { return E{~0ull}; }
};
template <> struct SortKey<AXB_F64> {
    using E = E16;
    using T = unsigned long long;
    static __device__ __forceinline__ E make(T raw, uint64_t pos, bool desc) {
        uint64_t k = image64f(raw == 0x8000000000000000ull ? 0ull : raw);
        return E{desc ? ~k : k, pos};
    }
    static __device__ __forceinline__ int64_t index(const E &e) { return (int64_t)e.idx; }
    static __device__ __forceinline__ E sentinel() { return E{~0ull, ~0ull}; }
};
template <> struct SortKey<AXB_I64> {
    using E = E16;
    using T = unsigned long long;
    static __device__ __forceinline__ E make(T raw, uint64_t pos, bool desc) {
        uint64_t k = raw ^ 0x8000000000000000ull;
        return E{desc ? ~k : k, pos};
    }
    static __device__ __forceinline__ int64_t index(const E &e) { return (int64_t)e.idx; }
    static __device__ __forceinline__ E sentinel() { return E{~0ull, ~0ull}; }
};

constexpr int kSortThreads = 512;
constexpr int kTile = 2048; // elements of one shared-memory tile (E16: 32 KiB)

struct SortParams {
    const char *in;   // contiguous (outer, n, inner)
    char *values;     // contiguous (outer, k_out, inner) or nullptr
    int64_t *indices; // same shape or nullptr
    int64_t slices, n, inner, n_pad, k_out;
    int32_t desc;
};

// compare-exchange passes on a tile held in shared memory: sizes size_lo..size_hi (doubling), for every size the strides
// min(size, tile) / 2 .. 1. `pos0` is the tile's first position inside its (padded) slice: the direction bit comes from the
// position in the slice, not in the tile.
template <typename E> __device__ __forceinline__ void tile_passes(E *sh, int tile, int64_t pos0, int64_t size_lo, int64_t size_hi) {
    for (int64_t size = size_lo; size <= size_hi; size <<= 1) {
        for (int stride = (int)std::min<int64_t>(size >> 1, tile >> 1); stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (tile >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const bool up = (((pos0 + lo) & size) == 0);
                const E a = sh[lo], b = sh[lo + stride];
                if (lt(b, a) == up) {
                    sh[lo] = b;
                    sh[lo + stride] = a;
                }
            }
            __syncthreads();
        }
    }
}

// (1) slices of at most kTile elements: load + sort + store in one kernel, one CTA per slice
template <int DT> __global__ void __launch_bounds__(kSortThreads) sort_small_kernel(const __grid_constant__ SortParams q) {
    using K = SortKey<DT>;
    using E = typename K::E;
    using T = typename K::T;
    __shared__ E sh[kTile];
    const int tile = (int)q.n_pad;
    for (int64_t sl = blockIdx.x; sl < q.slices; sl += gridDim.x) {
        const int64_t o = sl / q.inner, in = sl % q.inner;
        const T *src = reinterpret_cast<const T *>(q.in) + o * q.n * q.inner + in;
        for (int i = threadIdx.x; i < tile; i += blockDim.x)
            sh[i] = i < q.n ? K::make(src[(int64_t)i * q.inner], (uint64_t)i, q.desc != 0) : K::sentinel();
        __syncthreads();
        tile_passes<E>(sh, tile, 0, 2, tile);
        const int64_t obase = o * q.k_out * q.inner + in;
        for (int i = threadIdx.x; i < q.k_out; i += blockDim.x) {
            const int64_t pos = K::index(sh[i]);
            if (q.values) // the element itself (keys identify -0 with +0)
                reinterpret_cast<T *>(q.values)[obase + (int64_t)i * q.inner] = src[pos * q.inner];
            if (q.indices)
                q.indices[obase + (int64_t)i * q.inner] = pos;
        }
        __syncthreads();
    }
}

// (2) longer slices: scratch of slices * n_pad elements
template <int DT> __global__ void __launch_bounds__(256) sort_load_kernel(const __grid_constant__ SortParams q, typename SortKey<DT>::E *scratch) {
    using K = SortKey<DT>;
    using T = typename K::T;
    const int64_t total = q.slices * q.n_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t sl = i / q.n_pad, pos = i % q.n_pad;
        const int64_t o = sl / q.inner, in = sl % q.inner;
        scratch[i] = pos < q.n ? K::make(reinterpret_cast<const T *>(q.in)[(o * q.n + pos) * q.inner + in], (uint64_t)pos, q.desc != 0)
                               : K::sentinel();
    }
}
template <typename E> __global__ void __launch_bounds__(kSortThreads) sort_tile_kernel(E *scratch, int64_t n_tiles, int64_t n_pad, int64_t size_lo,
                                                                                       int64_t size_hi) {
    __shared__ E sh[kTile];
    for (int64_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        E *g = scratch + tl * kTile;
        for (int i = threadIdx.x; i < kTile; i += blockDim.x)
            sh[i] = g[i];
        __syncthreads();
        tile_passes<E>(sh, kTile, (tl * kTile) % n_pad, size_lo, size_hi);
        for (int i = threadIdx.x; i < kTile; i += blockDim.x)
            g[i] = sh[i];
        __syncthreads();
    }
}
template <typename E> __global__ void __launch_bounds__(256) sort_global_step_kernel(E *scratch, int64_t pairs, int64_t n_pad, int64_t size, int64_t stride) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = 2 * t - (t & (stride - 1)); // position in the concatenation of padded slices (n_pad is a power of two)
        const bool up = (((lo % n_pad) & size) == 0);
        const E a = scratch[lo], b = scratch[lo + stride];
        if (lt(b, a) == up) {
            scratch[lo] = b;
            scratch[lo + stride] = a;
        }
    }
}
template <int DT> __global__ void __launch_bounds__(256) sort_store_kernel(const __grid_constant__ SortParams q, const typename SortKey<DT>::E *scratch) {
    using K = SortKey<DT>;
    using T = typename K::T;
    const int64_t total = q.slices * q.k_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t sl = i / q.k_out, j = i % q.k_out;
        const int64_t o = sl / q.inner, in = sl % q.inner;
        const int64_t pos = K::index(scratch[sl * q.n_pad + j]);
        const int64_t dst = (o * q.k_out + j) * q.inner + in;
        if (q.values)
            reinterpret_cast<T *>(q.values)[dst] = reinterpret_cast<const T *>(q.in)[(o * q.n + pos) * q.inner + in];
        if (q.indices)
            q.indices[dst] = pos;
    }
}

template <int DT> axb_status run_sort(const SortParams &q, cudaStream_t s, DeviceCtx *ctx) {
    using E = typename SortKey<DT>::E;
    const int sms = ctx->sm_count;
    if (q.n_pad <= kTile) {
        const int blocks = (int)std::min<int64_t>(q.slices, (int64_t)sms * 8);
        sort_small_kernel<DT><<<blocks, kSortThreads, 0, s>>>(q);
        AXB_CHECK_LAUNCH();
        return AXB_OK;
    }
    void *scratch_v = nullptr;
    const int64_t total = q.slices * q.n_pad;
    AXB_TRY(axb_alloc((size_t)total * sizeof(E), s, &scratch_v));
    E *scratch = static_cast<E *>(scratch_v);
    auto grid = [&](int64_t work, int threads) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + threads - 1) / threads, (int64_t)sms * 16)); };
    axb_status st = AXB_OK;
    auto body = [&]() -> axb_status {
        sort_load_kernel<DT><<<grid(total, 256), 256, 0, s>>>(q, scratch);
        AXB_CHECK_LAUNCH();
        const int64_t n_tiles = total / kTile;
        const int tile_blocks = (int)std::min<int64_t>(n_tiles, (int64_t)sms * 8);
        sort_tile_kernel<E><<<tile_blocks, kSortThreads, 0, s>>>(scratch, n_tiles, q.n_pad, 2, kTile);
        AXB_CHECK_LAUNCH();
        for (int64_t size = 2 * kTile; size <= q.n_pad; size <<= 1) {
            for (int64_t stride = size >> 1; stride >= kTile; stride >>= 1) {
                sort_global_step_kernel<E><<<grid(total / 2, 256), 256, 0, s>>>(scratch, total / 2, q.n_pad, size, stride);
                AXB_CHECK_LAUNCH();
            }
            sort_tile_kernel<E><<<tile_blocks, kSortThreads, 0, s>>>(scratch, n_tiles, q.n_pad, size, size);
            AXB_CHECK_LAUNCH();
        }
        sort_store_kernel<DT><<<grid(q.slices * q.k_out, 256), 256, 0, s>>>(q, scratch);
        AXB_CHECK_LAUNCH();
        return AXB_OK;
    };
    st = body();
    axb_free(scratch_v, s);
    return st;
}

// ---- masked_select --------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 256, kSelPer = 8, kSelChunk = kSelThreads * kSelPer;

struct SelParams {
    const char *in, *mask;
    int64_t in_s[AXB_MAX_DIMS], mask_s[AXB_MAX_DIMS], shape[AXB_MAX_DIMS]; // broadcast strides in bytes (0 = broadcast)
    int32_t ndim, linear; // linear: both operands contiguous and un-broadcast -> offset = i * element size
    int32_t esize, msize;
    int64_t total;
};
__device__ __forceinline__ void sel_offsets(const SelParams &q, int64_t i, int64_t *io, int64_t *mo) {
    if (q.linear) {
        *io = i * q.esize;
        *mo = i * q.msize;
        return;
    }
    int64_t a = 0, b = 0;
    for (int d = q.ndim - 1; d >= 0; --d) {
        const int64_t c = i % q.shape[d];
        i /= q.shape[d];
        a += c * q.in_s[d];
        b += c * q.mask_s[d];
    }
    *io = a;
    *mo = b;
}
__device__ __forceinline__ int block_exclusive_scan(int v, int *total) { // kSelThreads threads
    __shared__ int warp_sums[kSelThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
            x += y;
    }
    if (lane == 31)
        warp_sums[w] = x;
    __syncthreads();
    int base = 0, all = 0;
    for (int k = 0; k < kSelThreads / 32; ++k) {
        const int sk = warp_sums[k];
        if (k < w)
            base += sk;
        all += sk;
    }
    __syncthreads();
    *total = all;
    return base + x - v;
}
__global__ void __launch_bounds__(kSelThreads) sel_count_kernel(const __grid_constant__ SelParams q, uint32_t *counts) {
    const int64_t first = ((int64_t)blockIdx.x * kSelThreads + threadIdx.x) * kSelPer;
    int c = 0;
    for (int k = 0; k < kSelPer; ++k) {
        const int64_t i = first + k;
        if (i < q.total) {
            int64_t io, mo;
            sel_offsets(q, i, &io, &mo);
            c += *reinterpret_cast<const uint8_t *>(q.mask + mo) != 0;
        }
    }
    int total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0)
        counts[blockIdx.x] = (uint32_t)total;
}
// one CTA: exclusive scan of the chunk counts; offsets[n] = grand total
__global__ void __launch_bounds__(1024) sel_scan_kernel(const uint32_t *counts, int64_t *offsets, int64_t n) {
    __shared__ int64_t part[1024];
    const int64_t per = (n + blockDim.x - 1) / blockDim.x;
    const int64_t b = threadIdx.x * per, e = std::min<int64_t>(b + per, n);
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i)
        s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) {
            const int64_t v = part[t];
            part[t] = run;
            run += v;
        }
        offsets[n] = run;
    }
    __syncthreads();
    int64_t run = part[threadIdx.x];
    for (int64_t i = b; i < e; ++i) {
        offsets[i] = run;
        run += counts[i];
    }
}
template <int ES> struct Bytes {
    unsigned char b[ES];
};
template <int ES> __global__ void __launch_bounds__(kSelThreads) sel_gather_kernel(const __grid_constant__ SelParams q, const int64_t *offsets, char *out) {
    const int64_t first = ((int64_t)blockIdx.x * kSelThreads + threadIdx.x) * kSelPer;
    int64_t ioff[kSelPer];
    unsigned flags = 0;
    int c = 0;
    for (int k = 0; k < kSelPer; ++k) {
        const int64_t i = first + k;
        ioff[k] = 0;
        if (i < q.total) {
            int64_t mo;
            sel_offsets(q, i, &ioff[k], &mo);
            if (*reinterpret_cast<const uint8_t *>(q.mask + mo) != 0) {
                flags |= 1u << k;
                ++c;
            }
        }
    }
    int total;
    int64_t dst = offsets[blockIdx.x] + block_exclusive_scan(c, &total);
    for (int k = 0; k < kSelPer; ++k)
        if (flags & (1u << k)) {
            if (ES == 1 || ES == 2 || ES == 4 || ES == 8 || ES == 16) {
                using V = typename std::conditional<ES == 1, uint8_t, typename std::conditional<ES == 2, uint16_t, typename std::conditional<ES == 4, uint32_t, typename std::conditional<ES == 8, uint64_t, uint4>::type>::type>::type>::type;
                reinterpret_cast<V *>(out)[dst] = *reinterpret_cast<const V *>(q.in + ioff[k]);
            }
            ++dst;
        }
}

} // namespace
} // namespace axb

using namespace axb;

extern "C" {

axb_status axb_pool(const axb_tensor *in, int is_avg, int spatial_dims, const int32_t *kernel, const int32_t *stride,
                    const int32_t *padding, int count_include_pad, const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_pool: input"));
    AXB_TRY(validate_tensor(out, "axb_pool: output"));
    AXB_REQUIRE(spatial_dims >= 1 && spatial_dims <= 3 && kernel && stride && padding, AXB_ERR_INVALID,
                "axb_pool: spatial_dims must be 1..3 with kernel / stride / padding arrays");
    AXB_REQUIRE(in->ndim == spatial_dims + 2 && out->ndim == in->ndim, AXB_ERR_SHAPE, "axb_pool: expected (N, C, spatial...) input and output");
    AXB_REQUIRE(is_float_dtype(in->dtype), AXB_ERR_UNSUPPORTED, "axb_pool: floating dtypes only (dispatch_float, operations.cpp:2387)");
    AXB_REQUIRE(out->dtype == in->dtype && is_contiguous(*out), AXB_ERR_INVALID, "axb_pool: output must be contiguous with the input's dtype");
    PoolParams q{};
    q.in = tensor_ptr(*in);
    q.out = tensor_ptr(*out);
    q.esize = (int32_t)dtype_size(in->dtype);
    q.include_pad = count_include_pad ? 1 : 0;
    q.in_s[0] = in->strides[0];
    q.in_s[1] = in->strides[1];
    q.C = in->shape[1];
    int64_t isz[3] = {1, 1, 1}, osz[3] = {1, 1, 1};
    for (int d = 0; d < 3; ++d) {
        q.k[d] = 1;
        q.s[d] = 1;
        q.p[d] = 0;
        q.in_s[2 + d] = 0;
    }
    AXB_REQUIRE(out->shape[0] == in->shape[0] && out->shape[1] == in->shape[1], AXB_ERR_SHAPE, "axb_pool: batch / channel extents differ");
    for (int d = 0; d < spatial_dims; ++d) {
        const int slot = 3 - spatial_dims + d; // 1-D pools use the W slot, 2-D the (H, W) slots
        AXB_REQUIRE(kernel[d] > 0 && stride[d] > 0 && padding[d] >= 0, AXB_ERR_INVALID, "axb_pool: kernel and stride must be positive, padding non-negative");
        q.k[slot] = kernel[d];
        q.s[slot] = stride[d];
        q.p[slot] = padding[d];
        q.in_s[2 + slot] = in->strides[2 + d];
        isz[slot] = in->shape[2 + d];
        const int64_t span = isz[slot] + 2 * (int64_t)padding[d] - kernel[d];
        AXB_REQUIRE(span >= 0, AXB_ERR_SHAPE, "axb_pool: window larger than the padded input");
        osz[slot] = span / stride[d] + 1; // pool_output_size, operations.cpp:2360-2363
        AXB_REQUIRE(out->shape[2 + d] == osz[slot], AXB_ERR_SHAPE, "axb_pool: output extent %lld of spatial dim %d, expected %lld",
                    (long long)out->shape[2 + d], d, (long long)osz[slot]);
    }
    q.D = isz[0]; q.H = isz[1]; q.W = isz[2];
    q.OD = osz[0]; q.OH = osz[1]; q.OW = osz[2];
    q.total = numel(*out);
    if (q.total == 0)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    switch (in->dtype) {
    case AXB_F32: return launch_pool<AXB_F32>(q, is_avg != 0, s, ctx->sm_count);
    case AXB_F64: return launch_pool<AXB_F64>(q, is_avg != 0, s, ctx->sm_count);
    case AXB_BF16: return launch_pool<AXB_BF16>(q, is_avg != 0, s, ctx->sm_count);
    default: return launch_pool<AXB_F16>(q, is_avg != 0, s, ctx->sm_count);
    }
}

axb_status axb_sort(const axb_tensor *in, int axis, int descending, const axb_tensor *values, const axb_tensor *indices,
                    axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_sort: input"));
    AXB_REQUIRE(values || indices, AXB_ERR_INVALID, "axb_sort: neither values nor indices requested");
    AXB_REQUIRE(in->dtype == AXB_F32 || in->dtype == AXB_F64 || in->dtype == AXB_I32 || in->dtype == AXB_I64, AXB_ERR_UNSUPPORTED,
                "axb_sort: keys must be float32 / float64 / int32 / int64 (the reference converts every other dtype to float32 "
                "first, operations.cpp:4199-4257)");
    AXB_REQUIRE(in->ndim >= 1, AXB_ERR_SHAPE, "axb_sort: rank-0 input");
    if (axis < 0)
        axis += in->ndim;
    AXB_REQUIRE(axis >= 0 && axis < in->ndim, AXB_ERR_INVALID, "axb_sort: axis out of range");
    AXB_REQUIRE(is_contiguous(*in), AXB_ERR_INVALID, "axb_sort: input must be contiguous (the reference gathers first, operations.cpp:4200)");
    const axb_tensor *any = values ? values : indices;
    const int64_t k_out = any->ndim == in->ndim ? any->shape[axis] : -1;
    for (const axb_tensor *o : {values, indices}) {
        if (!o)
            continue;
        AXB_TRY(validate_tensor(o, "axb_sort: output"));
        AXB_REQUIRE(o->ndim == in->ndim && is_contiguous(*o), AXB_ERR_SHAPE, "axb_sort: outputs must be contiguous with the input's rank");
        for (int d = 0; d < in->ndim; ++d)
            AXB_REQUIRE(d == axis ? o->shape[d] == k_out : o->shape[d] == in->shape[d], AXB_ERR_SHAPE, "axb_sort: output extent mismatch in dim %d", d);
    }
    AXB_REQUIRE(!values || values->dtype == in->dtype, AXB_ERR_UNSUPPORTED, "axb_sort: values must have the input's dtype");
    AXB_REQUIRE(!indices || indices->dtype == AXB_I64, AXB_ERR_UNSUPPORTED, "axb_sort: indices must be int64");
    SortParams q{};
    q.n = in->shape[axis];
    AXB_REQUIRE(k_out >= 0 && k_out <= q.n, AXB_ERR_SHAPE, "axb_sort: k = %lld outside 0..%lld", (long long)k_out, (long long)q.n);
    q.in = tensor_ptr(*in);
    q.values = values ? tensor_ptr(*values) : nullptr;
    q.indices = indices ? reinterpret_cast<int64_t *>(tensor_ptr(*indices)) : nullptr;
    q.inner = 1;
    int64_t outer = 1;
    for (int d = 0; d < axis; ++d)
        outer *= in->shape[d];
    for (int d = axis + 1; d < in->ndim; ++d)
        q.inner *= in->shape[d];
    q.slices = outer * q.inner;
    q.k_out = k_out;
    q.desc = descending ? 1 : 0;
    if (q.slices == 0 || q.n == 0 || k_out == 0)
        return AXB_OK;
    AXB_REQUIRE(q.n < (1ll << 31), AXB_ERR_UNSUPPORTED, "axb_sort: more than 2^31 elements along the sorted axis");
    q.n_pad = 2;
    while (q.n_pad < q.n)
        q.n_pad <<= 1;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    switch (in->dtype) {
    case AXB_F32: return run_sort<AXB_F32>(q, s, ctx);
    case AXB_F64: return run_sort<AXB_F64>(q, s, ctx);
    case AXB_I32: return run_sort<AXB_I32>(q, s, ctx);
    default: return run_sort<AXB_I64>(q, s, ctx);
    }
}

axb_status axb_masked_select(const axb_tensor *in, const axb_tensor *mask, void **out_data, int64_t *out_count, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_masked_select: input"));
    AXB_TRY(validate_tensor(mask, "axb_masked_select: mask"));
    AXB_REQUIRE(out_data && out_count, AXB_ERR_INVALID, "axb_masked_select: null result pointers");
    *out_data = nullptr;
    *out_count = 0;
    SelParams q{};
    int nd = 0;
    AXB_TRY(axb_broadcast_shape(in->shape, in->ndim, mask->shape, mask->ndim, q.shape, &nd));
    q.ndim = nd;
    q.total = 1;
    for (int d = 0; d < nd; ++d)
        q.total *= q.shape[d];
    if (q.total == 0)
        return AXB_OK;
    q.in = tensor_ptr(*in);
    q.mask = tensor_ptr(*mask);
    q.esize = (int32_t)dtype_size(in->dtype);
    q.msize = (int32_t)dtype_size(mask->dtype);
    for (int d = 0; d < nd; ++d) {
        const int di = d - (nd - in->ndim), dm = d - (nd - mask->ndim);
        q.in_s[d] = (di >= 0 && in->shape[di] != 1) ? in->strides[di] : 0;
        q.mask_s[d] = (dm >= 0 && mask->shape[dm] != 1) ? mask->strides[dm] : 0;
    }
    q.linear = (is_contiguous(*in) && is_contiguous(*mask) && numel(*in) == q.total && numel(*mask) == q.total) ? 1 : 0;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const int64_t chunks = (q.total + kSelChunk - 1) / kSelChunk;
    AXB_REQUIRE(chunks < (1ll << 31), AXB_ERR_UNSUPPORTED, "axb_masked_select: tensor too large");
    void *ws = nullptr;
    AXB_TRY(axb_alloc((size_t)chunks * 4 + (size_t)(chunks + 1) * 8 + 16, s, &ws));
    int64_t *offsets = static_cast<int64_t *>(ws);
    uint32_t *counts = reinterpret_cast<uint32_t *>(offsets + chunks + 1);
    void *result = nullptr;
    auto body = [&]() -> axb_status {
        sel_count_kernel<<<(unsigned)chunks, kSelThreads, 0, s>>>(q, counts);
        AXB_CHECK_LAUNCH();
        sel_scan_kernel<<<1, 1024, 0, s>>>(counts, offsets, chunks);
        AXB_CHECK_LAUNCH();
        int64_t count = 0;
        AXB_CUDA(cudaMemcpyAsync(&count, offsets + chunks, sizeof(count), cudaMemcpyDeviceToHost, s));
        AXB_CUDA(cudaStreamSynchronize(s));
        count_d2h(sizeof(count));
        *out_count = count;
        if (count == 0)
            return AXB_OK;
        AXB_TRY(axb_alloc((size_t)count * (size_t)q.esize, s, &result));
        char *o = static_cast<char *>(result);
        switch (q.esize) {
        case 1: sel_gather_kernel<1><<<(unsigned)chunks, kSelThreads, 0, s>>>(q, offsets, o); break;
        case 2: sel_gather_kernel<2><<<(unsigned)chunks, kSelThreads, 0, s>>>(q, offsets, o); break;
        case 4: sel_gather_kernel<4><<<(unsigned)chunks, kSelThreads, 0, s>>>(q, offsets, o); break;
        case 8: sel_gather_kernel<8><<<(unsigned)chunks, kSelThreads, 0, s>>>(q, offsets, o); break;
        default: sel_gather_kernel<16><<<(unsigned)chunks, kSelThreads, 0, s>>>(q, offsets, o); break;
        }
        AXB_CHECK_LAUNCH();
        return AXB_OK;
    };
    const axb_status st = body();
    axb_free(ws, s);
    if (st != AXB_OK) {
        if (result)
            axb_free(result, s);
        *out_count = 0;
        return st;
    }
    *out_data = result;
    return AXB_OK;
}

} // extern "C"
