// attention.cu — scaled dot-product attention as ONE C-ABI call (SURVEY.md §8 "next" row 1).
//
// Replaces ops::scaled_dot_product_attention -> cpu_flash_attention (src/tensor/operations.cpp:3979-4044,
// src/backends/cpu/flash_attention_cpu.cpp:18-131) behind the operations.cpp:4031 hook. fp32 in / fp32 out, the reference's
// semantics: S = Q K^T * scale, causal masks kv > q, an explicit mask marks positions to DROP (non-zero = masked out,
// flash_attention_cpu.cpp:91-105), softmax over kv, O = P V.
//
// Three launches instead of the six of the composed form (scale, bmm, masked_fill, softmax, split, bmm):
//   1. S = Q K^T             axb_matmul (fp32 on the tensor cores through the exact bf16 x 3 split, gemm_tcgen05.cu)
//   2. attn_softmax_kernel   one CTA per score row: scale, causal / explicit mask, max, exp, sum and normalisation in
//                            registers, and the probabilities leave DIRECTLY in the three-piece bf16 layout the next
//                            GEMM consumes — P is never written or re-read as fp32 (at whisper's 48 x 1500 x 1500 that
//                            is 432 MB less written and 432 MB less read per layer, and no separate split pass)
//   3. O = P V               the same GEMM with its A operand marked as pre-split
// The score matrix itself is still materialised once (fp32); a single-kernel flash attention with the split done in
// shared memory is the follow-up. Small problems (tests) take the same kernels with P written as fp32 in place of S.
#include <algorithm>
#include <cstring>

#include "axb_internal.h"
#include "gemm.h"
#include "reduce_common.cuh"

namespace axb {
namespace fa { // flash_attention.cu: the single-kernel path
bool supported(int32_t dtype, int64_t D, int64_t Dv, int64_t BH, int64_t Lq, int64_t Lk);
axb_status run(const axb_tensor *q, const axb_tensor *k, const axb_tensor *v, const uint8_t *mask, const int64_t mask_strides[4],
               float scale, int is_causal, const axb_tensor *out, DeviceCtx *ctx, cudaStream_t s);
} // namespace fa
namespace {

struct AttnParams {
    float *s;            // scores, contiguous [rows][lk]; overwritten with P when pieces == nullptr
    uint16_t *pieces;    // [rows][3 * kp] bf16 (hi | mid | lo), or nullptr
    const uint8_t *mask; // optional, one byte per element
    int64_t m_sb, m_sh, m_sq, m_sk; // mask byte strides over (batch, head, q, kv); 0 where broadcast
    int64_t rows, lq, lk, kp, heads;
    float scale;
    int32_t causal;
};

__device__ __forceinline__ float attn_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void attn_split3(float x, uint16_t &hi, uint16_t &mid, uint16_t &lo) {
    hi = red::f2bf(x);
    const float r1 = x - red::bf2f(hi);
    mid = red::f2bf(r1);
    lo = red::f2bf(r1 - red::bf2f(mid));
}
template <class F> __device__ __forceinline__ float block_allreduce_256(float v, float *smem, F f) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        v = f(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    __syncthreads(); // scratch reuse between consecutive reductions
    if ((threadIdx.x & 31) == 0)
        smem[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = smem[threadIdx.x & 7];
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1)
        t = f(t, __shfl_xor_sync(0xFFFFFFFFu, t, d));
    return t;
}

// One 256-thread CTA per score row. RESIDENT: the row (lk <= 256 * VPT) lives in registers, every thread owns runs of 8
// consecutive kv positions: two 128-bit loads of S and one 128-bit store per bf16 piece. Otherwise the row is re-read.
template <int VPT, bool RESIDENT> __global__ void __launch_bounds__(256) attn_softmax_kernel(const __grid_constant__ AttnParams p) {
    __shared__ float scratch[8];
    const float L2E = 1.4426950408889634f;
    constexpr int NCH = RESIDENT ? VPT / 8 : 1;
    // 128-bit accesses need 16-byte aligned rows: lk % 4 == 0 for S (fp32) and kp % 64 == 0 for the pieces (always true)
    const bool vec = RESIDENT && (p.lk & 3) == 0 && (reinterpret_cast<uintptr_t>(p.s) & 15) == 0;
    for (int64_t row = blockIdx.x; row < p.rows; row += gridDim.x) {
        const int64_t q = row % p.lq, bh = row / p.lq, h = bh % p.heads, b = bh / p.heads;
        float *srow = p.s + row * p.lk;
        const uint8_t *mrow = p.mask ? p.mask + b * p.m_sb + h * p.m_sh + q * p.m_sq : nullptr;
        const int64_t kmax = p.causal ? min(p.lk, q + 1) : p.lk; // kv positions > q are masked out
        auto masked = [&](int64_t k) { return k >= kmax || (mrow && mrow[k * p.m_sk]); };
        auto load = [&](int64_t k) -> float { return masked(k) ? -INFINITY : srow[k] * p.scale; };
        float x[RESIDENT ? VPT : 1];
        float m = -INFINITY;
        if (RESIDENT) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int64_t k0 = ((int64_t)c * 256 + threadIdx.x) * 8;
                if (vec && k0 + 8 <= p.lk) {
                    const float4 a = *reinterpret_cast<const float4 *>(srow + k0), d = *reinterpret_cast<const float4 *>(srow + k0 + 4);
                    const float v[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        x[c * 8 + i] = masked(k0 + i) ? -INFINITY : v[i] * p.scale;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        x[c * 8 + i] = k0 + i < p.lk ? load(k0 + i) : -INFINITY;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    m = fmaxf(m, x[c * 8 + i]);
            }
        } else {
            for (int64_t k = threadIdx.x; k < p.lk; k += 256)
                m = fmaxf(m, load(k));
        }
        m = block_allreduce_256(m, scratch, [](float a, float c) { return fmaxf(a, c); });
        // exp(x - m) as 2^(x * log2 e - m * log2 e): one FMA + MUFU.EX2; the common factor's rounding cancels in e / sum
        const float ml = m * L2E;
        float s = 0.0f;
        if (RESIDENT) {
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                x[i] = attn_ex2(fmaf(x[i], L2E, -ml));
                s += x[i];
            }
        } else {
            for (int64_t k = threadIdx.x; k < p.lk; k += 256)
                s += attn_ex2(fmaf(load(k), L2E, -ml));
        }
        s = block_allreduce_256(s, scratch, [](float a, float c) { return a + c; });
        const float inv = __fdiv_rn(1.0f, s); // fully masked row: m = -inf -> NaN like the reference's x - max
        uint16_t *prow = p.pieces ? p.pieces + row * 3 * p.kp : nullptr;
        auto emit = [&](int64_t k, float e) {
            const float y = e * inv;
            if (prow) {
                uint16_t hi, mid, lo;
                attn_split3(y, hi, mid, lo);
                prow[k] = hi;
                prow[p.kp + k] = mid;
                prow[2 * p.kp + k] = lo;
            } else {
                srow[k] = y;
            }
        };
        if (RESIDENT) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int64_t k0 = ((int64_t)c * 256 + threadIdx.x) * 8;
                if (k0 >= p.kp)
                    continue;
                if (prow && (k0 + 8 <= p.lk || k0 >= p.lk)) { // a full run (or pure K padding): 3 x 128-bit stores
                    uint32_t hh[4], mm[4], ll[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        red::split3_pair(k0 < p.lk ? x[c * 8 + 2 * i] * inv : 0.0f, k0 < p.lk ? x[c * 8 + 2 * i + 1] * inv : 0.0f,
                                         hh[i], mm[i], ll[i]);
                    *reinterpret_cast<uint4 *>(prow + k0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<uint4 *>(prow + p.kp + k0) = make_uint4(mm[0], mm[1], mm[2], mm[3]);
                    *reinterpret_cast<uint4 *>(prow + 2 * p.kp + k0) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (k0 + i < p.lk) {
                            emit(k0 + i, x[c * 8 + i]);
                        } else if (prow && k0 + i < p.kp) { // K padding inside a ragged run
                            prow[k0 + i] = 0;
                            prow[p.kp + k0 + i] = 0;
                            prow[2 * p.kp + k0 + i] = 0;
                        }
                    }
                }
            }
            if (prow) // K padding beyond the runs the resident loop covers
                for (int64_t k = (int64_t)NCH * 2048 + threadIdx.x; k < p.kp; k += 256) {
                    prow[k] = 0;
                    prow[p.kp + k] = 0;
                    prow[2 * p.kp + k] = 0;
                }
        } else {
            for (int64_t k = threadIdx.x; k < p.lk; k += 256)
                emit(k, attn_ex2(fmaf(load(k), L2E, -ml)));
            if (prow)
                for (int64_t k = p.lk + threadIdx.x; k < p.kp; k += 256) {
                    prow[k] = 0;
                    prow[p.kp + k] = 0;
                    prow[2 * p.kp + k] = 0;
                }
        }
    }
}

// Rows of up to 256 * RUNS / ... keys with ONE WARP per row (8 rows per CTA): each lane owns RUNS runs of 8 consecutive kv
// positions, the two row statistics need shuffles only — no block barrier at all (the 256-thread variant above spends two
// __syncthreads per 6 KB row; whisper's 1500-key rows: 295 -> see profiles). Pre-split output only.
template <int RUNS> __global__ void __launch_bounds__(256) attn_softmax_warp_kernel(const __grid_constant__ AttnParams p) {
    const float L2E = 1.4426950408889634f;
    const int lane = threadIdx.x & 31;
    const bool vec = (p.lk & 3) == 0 && (reinterpret_cast<uintptr_t>(p.s) & 15) == 0;
    for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < p.rows; row += (int64_t)gridDim.x * 8) {
        const int64_t q = row % p.lq, bh = row / p.lq, h = bh % p.heads, b = bh / p.heads;
        const float *srow = p.s + row * p.lk;
        const uint8_t *mrow = p.mask ? p.mask + b * p.m_sb + h * p.m_sh + q * p.m_sq : nullptr;
        const int64_t kmax = p.causal ? min(p.lk, q + 1) : p.lk;
        auto masked = [&](int64_t k) { return k >= kmax || (mrow && mrow[k * p.m_sk]); };
        float x[RUNS * 8];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < RUNS; ++c) {
            const int64_t k0 = ((int64_t)c * 32 + lane) * 8;
            if (vec && k0 + 8 <= p.lk) {
                const float4 a = *reinterpret_cast<const float4 *>(srow + k0), d = *reinterpret_cast<const float4 *>(srow + k0 + 4);
                const float v[8] = {a.x, a.y, a.z, a.w, d.x, d.y, d.z, d.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    x[c * 8 + i] = masked(k0 + i) ? -INFINITY : v[i] * p.scale;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    x[c * 8 + i] = (k0 + i < p.lk && !masked(k0 + i)) ? srow[k0 + i] * p.scale : -INFINITY;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                m = fmaxf(m, x[c * 8 + i]);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
            m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
        const float ml = m * L2E;
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < RUNS * 8; ++i) {
            x[i] = attn_ex2(fmaf(x[i], L2E, -ml));
            s += x[i];
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1)
            s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
        const float inv = __fdiv_rn(1.0f, s);
        uint16_t *prow = p.pieces + row * 3 * p.kp;
#pragma unroll
        for (int c = 0; c < RUNS; ++c) {
            const int64_t k0 = ((int64_t)c * 32 + lane) * 8;
            if (k0 >= p.kp)
                continue;
            uint32_t hh[4], mm[4], ll[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) // positions >= lk carry exp(-inf) = 0: the K padding comes out as zeros
                red::split3_pair(x[c * 8 + 2 * i] * inv, x[c * 8 + 2 * i + 1] * inv, hh[i], mm[i], ll[i]);
            *reinterpret_cast<uint4 *>(prow + k0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4 *>(prow + p.kp + k0) = make_uint4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<uint4 *>(prow + 2 * p.kp + k0) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
    }
}

} // namespace
} // namespace axb

using namespace axb;

static int &sdpa_last_path() {
    static thread_local int v = -1;
    return v;
}

extern "C" {

int axb_sdpa_last_path(void) { return sdpa_last_path(); }

axb_status axb_sdpa(const axb_tensor *q, const axb_tensor *k, const axb_tensor *v, const axb_tensor *mask, float scale,
                    int is_causal, const axb_tensor *out, axb_stream_t stream) {
    const char *what = "axb_sdpa";
    AXB_TRY(validate_tensor(q, what));
    AXB_TRY(validate_tensor(k, what));
    AXB_TRY(validate_tensor(v, what));
    AXB_TRY(validate_tensor(out, what));
    AXB_REQUIRE(q->ndim == 4 && k->ndim == 4 && v->ndim == 4 && out->ndim == 4, AXB_ERR_SHAPE, "%s: q, k, v, out must be 4-D (batch, heads, seq, dim)", what);
    AXB_REQUIRE((q->dtype == AXB_F32 || q->dtype == AXB_BF16 || q->dtype == AXB_F16) && k->dtype == q->dtype && v->dtype == q->dtype &&
                    out->dtype == q->dtype,
                AXB_ERR_UNSUPPORTED, "%s: q, k, v, out must share one of Float32 / BFloat16 / Float16", what);
    const int64_t B = q->shape[0], H = q->shape[1], Lq = q->shape[2], D = q->shape[3], Lk = k->shape[2], Dv = v->shape[3];
    AXB_REQUIRE(k->shape[0] == B && v->shape[0] == B && k->shape[1] == H && v->shape[1] == H, AXB_ERR_SHAPE, "%s: batch / heads mismatch", what);
    AXB_REQUIRE(k->shape[3] == D && v->shape[2] == Lk, AXB_ERR_SHAPE, "%s: head_dim / kv length mismatch", what);
    AXB_REQUIRE(out->shape[0] == B && out->shape[1] == H && out->shape[2] == Lq && out->shape[3] == Dv, AXB_ERR_SHAPE, "%s: out shape", what);
    if (numel(*out) == 0)
        return AXB_OK;
    AXB_REQUIRE(Lk > 0 && D > 0, AXB_ERR_SHAPE, "%s: empty key sequence", what);
    if (!(scale > 0.0f))
        scale = 1.0f / sqrtf((float)D); // operations.cpp:4013-4016
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));

    AttnParams p;
    memset(&p, 0, sizeof p);
    if (mask) {
        AXB_TRY(validate_tensor(mask, what));
        AXB_REQUIRE(dtype_size(mask->dtype) == 1, AXB_ERR_UNSUPPORTED, "%s: mask must be Bool / 8-bit (non-zero = masked out)", what);
        AXB_REQUIRE(mask->ndim <= 4, AXB_ERR_SHAPE, "%s: mask rank", what);
        const int64_t want[4] = {B, H, Lq, Lk};
        int64_t st[4] = {0, 0, 0, 0};
        for (int i = 0; i < mask->ndim; ++i) {
            const int fi = i + (4 - mask->ndim);
            AXB_REQUIRE(mask->shape[i] == want[fi] || mask->shape[i] == 1, AXB_ERR_SHAPE, "%s: mask does not broadcast to (B, H, Lq, Lk)", what);
            st[fi] = mask->shape[i] == 1 ? 0 : mask->strides[i];
        }
        p.mask = reinterpret_cast<const uint8_t *>(tensor_ptr(*mask));
        p.m_sb = st[0]; p.m_sh = st[1]; p.m_sq = st[2]; p.m_sk = st[3];
    }
    if (fa::supported(q->dtype, D, Dv, B * H, Lq, Lk)) { // one tcgen05 kernel, no materialised score matrix
        const int64_t mst[4] = {p.m_sb, p.m_sh, p.m_sq, p.m_sk};
        const axb_status fst = fa::run(q, k, v, p.mask, mst, scale, is_causal, out, ctx, s);
        if (fst == AXB_OK)
            sdpa_last_path() = 1;
        return fst;
    }
    AXB_REQUIRE(q->dtype == AXB_F32, AXB_ERR_UNSUPPORTED,
                "%s: head_dim > 64 is served by the three-kernel path, which is Float32 only (half types: matmul + softmax + matmul)", what);
    sdpa_last_path() = 0;
    const int64_t rows = B * H * Lq;
    void *sbuf = nullptr, *pbuf = nullptr;
    AXB_TRY(axb_alloc((size_t)rows * Lk * 4, s, &sbuf));
    const int64_t sshape[4] = {B, H, Lq, Lk};
    axb_tensor S = make_contiguous_desc(sbuf, AXB_F32, sshape, 4);
    axb_status st = axb_matmul(q, k, 0, 1, nullptr, &S, s); // S = Q K^T
    // the P V product runs with a pre-split A when it qualifies for the tensor-core path (see try_split_f32)
    const double flops = 2.0 * (double)Lq * (double)Dv * (double)Lk * (double)(B * H);
    const bool contiguous_out = is_contiguous(*out);
    const bool presplit = flops >= 1.0e8 && Lq >= 64 && Dv >= 64 && Lk >= 32 && B * H <= 65535 && contiguous_out &&
                          !getenv("AXB_SDPA_NO_PRESPLIT");
    const int64_t kp = (Lk + 63) / 64 * 64;
    if (st == AXB_OK && presplit)
        st = axb_alloc((size_t)rows * 3 * kp * 2, s, &pbuf);
    if (st == AXB_OK) {
        p.s = static_cast<float *>(sbuf);
        p.pieces = static_cast<uint16_t *>(pbuf);
        p.rows = rows; p.lq = Lq; p.lk = Lk; p.kp = kp; p.heads = H;
        p.scale = scale;
        p.causal = is_causal ? 1 : 0;
        const unsigned grid = (unsigned)std::min<int64_t>(rows, (int64_t)ctx->sm_count * 16);
        static const bool no_warp_rows = getenv("AXB_SDPA_NO_WARP_ROWS") != nullptr; // A/B switch
        if (presplit && kp <= 256 * 6 && !no_warp_rows)
            attn_softmax_warp_kernel<6><<<(unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->sm_count * 16), 256, 0, s>>>(p);
        else if (presplit && kp <= 256 * 8 && !no_warp_rows)
            attn_softmax_warp_kernel<8><<<(unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)ctx->sm_count * 16), 256, 0, s>>>(p);
        else if (Lk <= 256 * 8)
            attn_softmax_kernel<8, true><<<grid, 256, 0, s>>>(p);
        else if (Lk <= 256 * 32)
            attn_softmax_kernel<32, true><<<grid, 256, 0, s>>>(p);
        else
            attn_softmax_kernel<1, false><<<grid, 256, 0, s>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            st = fail(AXB_ERR_CUDA, "%s: softmax launch failed: %s", what, cudaGetErrorString(e));
        else
            count_launch();
    }
    if (st == AXB_OK) {
        if (presplit) {
            gemm::presplit_a_hint() = gemm::PresplitA{static_cast<const uint16_t *>(pbuf), kp};
            st = axb_matmul(&S, v, 0, 0, nullptr, out, s); // S only describes (B, H, Lq, Lk); the data comes from the pieces
            gemm::presplit_a_hint() = gemm::PresplitA{};
        } else {
            st = axb_matmul(&S, v, 0, 0, nullptr, out, s); // S now holds P
        }
    }
    axb_free(sbuf, s);
    if (pbuf)
        axb_free(pbuf, s);
    return st;
}

} // extern "C"
