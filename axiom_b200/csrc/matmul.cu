// matmul.cu — kernel family 3 front end: shape rules, batch broadcasting, stride handling and dispatch.
//
// Replaces CPUMatMulOperation::execute_matmul (src/backends/cpu/cpu_operations.cpp:2089-2150), matmul_2d (:1773-1883:
// zero-copy transposes via strides -> BLAS trans flags) and BlasBackend::sgemm (blas/blas_backend.hpp:38); shape rules
// infer_matmul_shape (src/graph/graph_registry.cpp:147-220); MatMulActivationStep (compiled_graph.hpp:54-58) and
// nn::Linear's bias add (src/nn/linear.cpp:22-25) are the fused epilogue.
//
// Two device paths:
//   * gemm_tcgen05.cu — bf16 / fp16 operands whose K- or MN-major layout is TMA-describable (16-byte aligned strides):
//     tcgen05.mma with TMEM accumulators fed by a TMA ring, fp32 accumulate, one rounding at the end. The C4 path.
//   * the SIMT kernel below — everything else (fp32, fp64, odd strides / alignments, tiny shapes): a 64x64x16 shared-
//     memory tile with 4x4 register blocking, FFMA in fp32 (fp64 for Float64), arbitrary element strides and batch
//     broadcasting. fp32 has no IEEE tensor-core MMA on sm_100a (SURVEY.md §7 hard parts); it is correct at sgemm
//     precision and fast enough for the README chain, the roofline target is stated for bf16.
#include <algorithm>
#include <cstring>

#include "axb_internal.h"
#include "gemm.h"
#include "hwy_math.cuh"
#include "reduce_common.cuh"

namespace axb {
namespace gemm {

using red::bf16_t;
using red::f16_t;

template <typename T> struct ComputeOf { typedef float type; };
template <> struct ComputeOf<double> { typedef double type; };

template <typename T, typename C> __device__ __forceinline__ C ld_elem(const char *p) { return (C) * reinterpret_cast<const T *>(p); }
template <> __device__ __forceinline__ float ld_elem<bf16_t, float>(const char *p) { return red::bf2f(*reinterpret_cast<const uint16_t *>(p)); }
template <> __device__ __forceinline__ float ld_elem<f16_t, float>(const char *p) { return red::h2f(*reinterpret_cast<const uint16_t *>(p)); }
template <typename T, typename C> __device__ __forceinline__ void st_elem(char *p, C v) { *reinterpret_cast<T *>(p) = (T)v; }
template <> __device__ __forceinline__ void st_elem<bf16_t, float>(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = red::f2bf(v); }
template <> __device__ __forceinline__ void st_elem<f16_t, float>(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = red::f2h(v); }

template <typename T, typename C> __device__ __forceinline__ C round_out(C v) { return v; }
template <> __device__ __forceinline__ float round_out<bf16_t, float>(float v) { return red::bf2f(red::f2bf(v)); }
template <> __device__ __forceinline__ float round_out<f16_t, float>(float v) { return red::h2f(red::f2h(v)); }

__device__ __forceinline__ float ld_any_f(const char *p, int dt) {
    switch (dt) {
    case AXB_F32: return *reinterpret_cast<const float *>(p);
    case AXB_BF16: return red::bf2f(*reinterpret_cast<const uint16_t *>(p));
    case AXB_F16: return red::h2f(*reinterpret_cast<const uint16_t *>(p));
    case AXB_F64: return (float)*reinterpret_cast<const double *>(p);
    default: return 0.0f;
    }
}

// activation of the fused epilogue, in the reference's fp32 formulas (hwy_kernels-inl.h:786-880)
__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
    case AXB_OP_RELU: return v > 0.0f ? v : 0.0f;
    case AXB_OP_GELU: return hwym::Gelu(v);
    case AXB_OP_SILU: return hwym::Silu(v);
    case AXB_OP_SIGMOID: return hwym::Sigmoid(v);
    case AXB_OP_TANH: return hwym::Tanh(v);
    default: return v;
    }
}

__device__ __forceinline__ void batch_offsets(const Desc &d, int64_t b, int64_t *ao, int64_t *bo, int64_t *co, int64_t *biaso) {
    int64_t a = 0, bb = 0, c = 0, bi = 0;
#pragma unroll 1
    for (int i = 0; i < d.nb; ++i) {
        int64_t q = b / d.bshape[i];
        int64_t k = b - q * d.bshape[i];
        b = q;
        a += k * d.a_bs[i];
        bb += k * d.b_bs[i];
        c += k * d.c_bs[i];
        bi += k * d.bias_bs[i];
    }
    *ao = a; *bo = bb; *co = c; *biaso = bi;
}

// 64x64 output tile per CTA, K step 16, 256 threads x (4x4) outputs.
template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __grid_constant__ Desc d) {
    typedef typename ComputeOf<TC>::type C;
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ C sa[BK][BM + 4];
    __shared__ C sb[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    int64_t ao, bo, co, biaso;
    batch_offsets(d, blockIdx.z, &ao, &bo, &co, &biaso);
    const char *A = d.a + ao, *B = d.b + bo;
    C acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j] = C(0);
    // loader mapping: pick the thread->element order that makes the faster-varying thread index walk the smaller stride
    const bool a_kfast = llabs(d.a_cs) <= llabs(d.a_rs); // A[m][k]: k contiguous?
    const bool b_nfast = llabs(d.b_cs) <= llabs(d.b_rs); // B[k][n]: n contiguous?
    for (int64_t k0 = 0; k0 < d.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256; // 0..1023 over a 64x16 tile
            int m, k;
            if (a_kfast) { k = e & 15; m = e >> 4; } else { m = e & 63; k = e >> 6; }
            C v = C(0);
            if (m0 + m < d.M && k0 + k < d.K)
                v = ld_elem<TA, C>(A + (m0 + m) * d.a_rs + (k0 + k) * d.a_cs);
            sa[k][m] = v;
            int n, kk;
            if (b_nfast) { n = e & 63; kk = e >> 6; } else { kk = e & 15; n = e >> 4; }
            C w = C(0);
            if (n0 + n < d.N && k0 + kk < d.K)
                w = ld_elem<TB, C>(B + (k0 + kk) * d.b_rs + (n0 + n) * d.b_cs);
            sb[kk][n] = w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            C ra[4], rb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                ra[i] = sa[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                rb[j] = sb[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += ra[i] * rb[j];
        }
        __syncthreads();
    }
    char *Cp = d.c + co;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= d.M)
            continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n >= d.N)
                continue;
            C v = acc[i][j];
            // the reference runs matmul, bias add and activation as separate ops: each result is rounded to the
            // output dtype before the next one (nn/linear.cpp:22-25, MatMulActivationStep)
            if (d.bias) {
                v = round_out<TC, C>(v);
                v = v + (C)ld_any_f(d.bias + biaso + m * d.bias_rs + n * d.bias_cs, d.bias_dtype);
            }
            if (d.act) {
                v = round_out<TC, C>(v);
                v = (C)apply_act((float)v, d.act);
            }
            st_elem<TC, C>(Cp + m * d.c_rs + n * d.c_cs, v);
        }
    }
}

namespace {

template <typename TA, typename TB, typename TC> axb_status launch_simt(const Desc &d, cudaStream_t s) {
    dim3 grid((unsigned)((d.N + 63) / 64), (unsigned)((d.M + 63) / 64), (unsigned)d.batch);
    AXB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, AXB_ERR_UNSUPPORTED, "axb_matmul: M or batch too large for the SIMT path");
    gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, s>>>(d);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

template <typename TA> axb_status simt_by_b(const Desc &d, int bdt, int cdt, cudaStream_t s);

axb_status simt_dispatch(const Desc &d, int adt, int bdt, int cdt, cudaStream_t s) {
#define AXB_GEMM_CASE(ADT, TA, BDT, TB, CDT, TC)                                                        \
    if (adt == ADT && bdt == BDT && cdt == CDT)                                                         \
        return launch_simt<TA, TB, TC>(d, s);
    AXB_GEMM_CASE(AXB_F32, float, AXB_F32, float, AXB_F32, float)
    AXB_GEMM_CASE(AXB_F64, double, AXB_F64, double, AXB_F64, double)
    AXB_GEMM_CASE(AXB_BF16, bf16_t, AXB_BF16, bf16_t, AXB_BF16, bf16_t)
    AXB_GEMM_CASE(AXB_F16, f16_t, AXB_F16, f16_t, AXB_F16, f16_t)
    // mixed operands promote (promote_types, operations.cpp:346-397)
    AXB_GEMM_CASE(AXB_BF16, bf16_t, AXB_F32, float, AXB_F32, float)
    AXB_GEMM_CASE(AXB_F32, float, AXB_BF16, bf16_t, AXB_F32, float)
    AXB_GEMM_CASE(AXB_F16, f16_t, AXB_F32, float, AXB_F32, float)
    AXB_GEMM_CASE(AXB_F32, float, AXB_F16, f16_t, AXB_F32, float)
    AXB_GEMM_CASE(AXB_F32, float, AXB_F64, double, AXB_F64, double)
    AXB_GEMM_CASE(AXB_F64, double, AXB_F32, float, AXB_F64, double)
#undef AXB_GEMM_CASE
    return fail(AXB_ERR_UNSUPPORTED, "axb_matmul: dtypes (%d, %d) -> %d have no GPU kernel (float types only)", adt, bdt, cdt);
}

struct Norm {
    int64_t bshape[AXB_MAX_DIMS];
    int nb;
    int64_t M, N, K;
    int64_t rs, cs;                  // byte strides of the (rows, cols) of the effective 2-D operand
    int64_t bstride[AXB_MAX_DIMS];   // byte strides of its own batch dims (right-aligned later)
    int own_nb;
};

} // namespace

// Normalises (a, b, flags) to a Desc; also yields the result shape. Pure host logic (no device needed).
axb_status build_desc(const axb_tensor *a, const axb_tensor *b, int ta, int tb, Desc *d, int64_t *oshape, int *ondim) {
    AXB_TRY(validate_tensor(a, "axb_matmul(a)"));
    AXB_TRY(validate_tensor(b, "axb_matmul(b)"));
    AXB_REQUIRE(a->ndim >= 1 && b->ndim >= 1, AXB_ERR_SHAPE, "axb_matmul: operands must have at least 1 dim");
    memset(d, 0, sizeof *d);
    // 1-D promotion: a -> {1,K}, b -> {K,1} (graph_registry.cpp:153-158)
    int64_t ash[AXB_MAX_DIMS + 1], ast[AXB_MAX_DIMS + 1], bsh[AXB_MAX_DIMS + 1], bst[AXB_MAX_DIMS + 1];
    int an = a->ndim, bn = b->ndim;
    if (an == 1) {
        ash[0] = 1; ast[0] = 0; ash[1] = a->shape[0]; ast[1] = a->strides[0];
        an = 2;
    } else
        for (int i = 0; i < an; ++i) { ash[i] = a->shape[i]; ast[i] = a->strides[i]; }
    if (bn == 1) {
        bsh[0] = b->shape[0]; bst[0] = b->strides[0]; bsh[1] = 1; bst[1] = 0;
        bn = 2;
    } else
        for (int i = 0; i < bn; ++i) { bsh[i] = b->shape[i]; bst[i] = b->strides[i]; }
    int64_t a_rows = ash[an - 2], a_cols = ash[an - 1], a_rs = ast[an - 2], a_cs = ast[an - 1];
    if (ta) { std::swap(a_rows, a_cols); std::swap(a_rs, a_cs); }
    int64_t b_rows = bsh[bn - 2], b_cols = bsh[bn - 1], b_rs = bst[bn - 2], b_cs = bst[bn - 1];
    if (tb) { std::swap(b_rows, b_cols); std::swap(b_rs, b_cs); }
    AXB_REQUIRE(a_cols == b_rows, AXB_ERR_SHAPE, "axb_matmul: inner dimensions differ (%lld vs %lld)", (long long)a_cols,
                (long long)b_rows);
    d->M = a_rows; d->N = b_cols; d->K = a_cols;
    d->a_rs = a_rs; d->a_cs = a_cs; d->b_rs = b_rs; d->b_cs = b_cs;
    // batch broadcasting (right-aligned); Desc batch dims are innermost-first
    const int anb = an - 2, bnb = bn - 2, nb = std::max(anb, bnb);
    AXB_REQUIRE(nb <= kMaxBatchDims, AXB_ERR_UNSUPPORTED, "axb_matmul: too many batch dims");
    int64_t full[AXB_MAX_DIMS];
    for (int i = 0; i < nb; ++i) {
        int ai = i - (nb - anb), bi = i - (nb - bnb);
        int64_t ad = ai >= 0 ? ash[ai] : 1, bd = bi >= 0 ? bsh[bi] : 1;
        AXB_REQUIRE(ad == bd || ad == 1 || bd == 1, AXB_ERR_SHAPE, "axb_matmul: batch dimensions do not broadcast (%lld vs %lld)",
                    (long long)ad, (long long)bd);
        full[i] = ad == 1 ? bd : ad;
        int slot = nb - 1 - i;
        d->bshape[slot] = full[i];
        d->a_bs[slot] = (ai >= 0 && ad != 1) ? ast[ai] : 0;
        d->b_bs[slot] = (bi >= 0 && bd != 1) ? bst[bi] : 0;
    }
    d->nb = nb;
    d->batch = 1;
    for (int i = 0; i < nb; ++i)
        d->batch *= full[i];
    // result shape, with the 1-D squeezes (graph_registry.cpp:203-214)
    int n = 0;
    for (int i = 0; i < nb; ++i)
        oshape[n++] = full[i];
    if (a->ndim != 1)
        oshape[n++] = d->M;
    if (b->ndim != 1)
        oshape[n++] = d->N;
    *ondim = n;
    d->a = tensor_ptr(*a);
    d->b = tensor_ptr(*b);
    return AXB_OK;
}

} // namespace gemm
} // namespace axb

using namespace axb;
using namespace axb::gemm;

extern "C" {

axb_status axb_matmul_shape(const axb_tensor *a, const axb_tensor *b, int transpose_a, int transpose_b, int64_t *out_shape,
                            int *out_ndim) {
    AXB_REQUIRE(out_shape && out_ndim, AXB_ERR_INVALID, "axb_matmul_shape: null output");
    Desc d;
    // shape inference must not need device pointers: patch null bases for the validation inside build_desc
    axb_tensor a2 = *a, b2 = *b;
    if (!a2.base)
        a2.base = reinterpret_cast<void *>(16);
    if (!b2.base)
        b2.base = reinterpret_cast<void *>(16);
    return build_desc(&a2, &b2, transpose_a, transpose_b, &d, out_shape, out_ndim);
}

axb_status axb_matmul(const axb_tensor *a, const axb_tensor *b, int transpose_a, int transpose_b,
                      const axb_matmul_epilogue *epilogue, const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(out, "axb_matmul(out)"));
    Desc d;
    int64_t oshape[AXB_MAX_DIMS];
    int ondim = 0;
    AXB_TRY(build_desc(a, b, transpose_a, transpose_b, &d, oshape, &ondim));
    const int32_t cdt = axb_promote_types(a->dtype, b->dtype);
    AXB_REQUIRE(out->dtype == cdt, AXB_ERR_INVALID, "axb_matmul: out dtype %d, expected promote(a,b) = %d", out->dtype, cdt);
    AXB_REQUIRE(out->ndim == ondim, AXB_ERR_SHAPE, "axb_matmul: out rank %d, expected %d", out->ndim, ondim);
    for (int i = 0; i < ondim; ++i)
        AXB_REQUIRE(out->shape[i] == oshape[i], AXB_ERR_SHAPE, "axb_matmul: out dim %d is %lld, expected %lld", i,
                    (long long)out->shape[i], (long long)oshape[i]);
    if (numel(*out) == 0)
        return AXB_OK;
    // output strides: rows/cols of the (possibly squeezed) result
    {
        int pos = ondim;
        d.c_cs = 0; d.c_rs = 0;
        if (b->ndim != 1) { --pos; d.c_cs = out->strides[pos]; }
        if (a->ndim != 1) { --pos; d.c_rs = out->strides[pos]; }
        for (int i = 0; i < d.nb; ++i)
            d.c_bs[i] = out->strides[pos - 1 - i];
    }
    d.c = tensor_ptr(*out);
    if (d.K == 0) { // empty contraction: zeros
        return axb_fill(out, 0.0, stream);
    }
    if (epilogue) {
        d.const_ops = epilogue->flags & (AXB_MM_CONST_A | AXB_MM_CONST_B);
        d.act = epilogue->activation;
        AXB_REQUIRE(d.act == 0 || d.act == AXB_OP_RELU || d.act == AXB_OP_GELU || d.act == AXB_OP_SILU ||
                        d.act == AXB_OP_SIGMOID || d.act == AXB_OP_TANH,
                    AXB_ERR_UNSUPPORTED, "axb_matmul: activation %d cannot be fused", d.act);
        if (epilogue->bias) {
            const axb_tensor *bt = epilogue->bias;
            AXB_TRY(validate_tensor(bt, "axb_matmul(bias)"));
            AXB_REQUIRE(bt->ndim <= d.nb + 2, AXB_ERR_SHAPE, "axb_matmul: bias rank too high");
            AXB_REQUIRE(a->ndim != 1 && b->ndim != 1, AXB_ERR_UNSUPPORTED, "axb_matmul: bias with 1-D operands is not supported");
            d.bias = tensor_ptr(*bt);
            d.bias_dtype = bt->dtype;
            const int full_nd = d.nb + 2;
            for (int i = 0; i < bt->ndim; ++i) {
                int fi = i + (full_nd - bt->ndim); // position in [batch..., M, N]
                int64_t want = fi == full_nd - 1 ? d.N : fi == full_nd - 2 ? d.M : d.bshape[d.nb - 1 - fi];
                AXB_REQUIRE(bt->shape[i] == want || bt->shape[i] == 1, AXB_ERR_SHAPE, "axb_matmul: bias does not broadcast");
                int64_t st = bt->shape[i] == 1 ? 0 : bt->strides[i];
                if (fi == full_nd - 1)
                    d.bias_cs = st;
                else if (fi == full_nd - 2)
                    d.bias_rs = st;
                else
                    d.bias_bs[d.nb - 1 - fi] = st;
            }
        }
    }
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    bool done = false;
    AXB_TRY(try_tcgen05(d, a->dtype, b->dtype, cdt, ctx, s, &done));
    if (done)
        return AXB_OK;
    last_path() = 0;
    return simt_dispatch(d, a->dtype, b->dtype, cdt, s);
}

int axb_matmul_last_path(void) { return axb::gemm::last_path(); }
axb_status axb_invalidate_cached(const void *ptr, size_t bytes) {
    axb::gemm::invalidate_const_cache(ptr, bytes);
    return AXB_OK;
}
axb_status axb_get_const_cache_stats(axb_const_cache_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_const_cache_stats: null out");
    *out = axb::gemm::const_cache_stats();
    return AXB_OK;
}

} // extern "C"
