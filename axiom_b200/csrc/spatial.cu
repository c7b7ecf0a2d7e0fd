// spatial.cu — conv2d, transposed convolution (1-D / 2-D) and interpolate: the spatial operators next to the named path
// ("remaining GPU op surface"). The reference routes them through MPSGraph on Metal
// (src/backends/metal/mpsgraph_operations.mm:2547-2760) and otherwise runs scalar CPU loops
// (src/tensor/operations.cpp:3141-3330 conv2d, :3340-3690 conv_transpose1d/2d, :3694-3975 interpolate).
//
// All three are gather-form kernels, one thread per OUTPUT element, consecutive threads on consecutive positions of the
// innermost output dim (coalesced stores, unit- or stride-s reads), inputs and weights addressed through their byte
// strides (views need no copy). Numerics mirror the CPU loops term by term so that fp32 results are bit-identical:
//   conv2d            fp64 accumulator over (ic, fh, fw), + bias in fp64, one rounding (operations.cpp:3263-3318)
//   conv_transpose    the reference SCATTERS "out += T(double(in) * double(w))" over (ic, ih, iw): for one output that is
//                     a sum in the element type T, ordered by ic, then ih, iw ascending — the gather form walks fh, fw
//                     DESCENDING to visit the same terms in the same order; bias added last in T (:3470-3486)
//   interpolate       source coordinates and weights in fp32 for every dtype (:3779-3955); nearest / (bi)linear / bicubic
#include <algorithm>
#include <cstring>

#include "axb_internal.h"
#include "gemm.h"
#include "reduce_common.cuh"

namespace axb {
namespace {

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ double ld(const char *p) { return (double)*reinterpret_cast<const float *>(p); }
    static __device__ __forceinline__ float ldf(const char *p) { return *reinterpret_cast<const float *>(p); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<float *>(p) = (float)v; }
    static __device__ __forceinline__ void stf(char *p, float v) { *reinterpret_cast<float *>(p) = v; }
    static __device__ __forceinline__ double round(double v) { return (double)(float)v; }
};
template <> struct Elem<double> {
    static __device__ __forceinline__ double ld(const char *p) { return *reinterpret_cast<const double *>(p); }
    static __device__ __forceinline__ float ldf(const char *p) { return (float)*reinterpret_cast<const double *>(p); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<double *>(p) = v; }
    static __device__ __forceinline__ void stf(char *p, float v) { *reinterpret_cast<double *>(p) = (double)v; }
    static __device__ __forceinline__ double round(double v) { return v; }
};
struct BF16 {};
struct F16 {};
template <> struct Elem<BF16> { // double -> float -> bf16 (bfloat16.hpp:16-27: RNE, NaN -> 0x7FC0)
    static __device__ __forceinline__ double ld(const char *p) { return (double)red::bf2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ float ldf(const char *p) { return red::bf2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<uint16_t *>(p) = red::f2bf((float)v); }
    static __device__ __forceinline__ void stf(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = red::f2bf(v); }
    static __device__ __forceinline__ double round(double v) { return (double)red::bf2f(red::f2bf((float)v)); }
};
template <> struct Elem<F16> {
    static __device__ __forceinline__ double ld(const char *p) { return (double)red::h2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ float ldf(const char *p) { return red::h2f(*reinterpret_cast<const uint16_t *>(p)); }
    static __device__ __forceinline__ void st(char *p, double v) { *reinterpret_cast<uint16_t *>(p) = red::f2h((float)v); }
    static __device__ __forceinline__ void stf(char *p, float v) { *reinterpret_cast<uint16_t *>(p) = red::f2h(v); }
    static __device__ __forceinline__ double round(double v) { return (double)red::h2f(red::f2h((float)v)); }
};

// ---- convolution (forward and transposed) ----------------------------------------------------------------------------
struct ConvParams {
    const char *in, *w, *bias;
    char *out;
    int64_t in_s[4], w_s[4], bias_s; // byte strides: input (N, C, H, W), weight (dim0, dim1, KH, KW)
    int64_t N, Cin, H, W, Cout, KH, KW, OH, OW;
    int64_t cin_g, cout_g;           // channels per group
    int32_t sh, sw, ph, pw, dh, dw, esize;
    int64_t total;
};

template <typename T> __global__ void __launch_bounds__(256) conv2d_kernel(const __grid_constant__ ConvParams p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ow = i % p.OW, t1 = i / p.OW, oh = t1 % p.OH, t2 = t1 / p.OH, oc = t2 % p.Cout, n = t2 / p.Cout;
        const int64_t g = oc / p.cout_g;
        double acc = 0.0;
        for (int64_t ic = 0; ic < p.cin_g; ++ic) {
            const char *ip = p.in + n * p.in_s[0] + (g * p.cin_g + ic) * p.in_s[1];
            const char *wp = p.w + oc * p.w_s[0] + ic * p.w_s[1];
            for (int64_t fh = 0; fh < p.KH; ++fh) {
                const int64_t ih = oh * p.sh - p.ph + fh * p.dh;
                if (ih < 0 || ih >= p.H)
                    continue;
                for (int64_t fw = 0; fw < p.KW; ++fw) {
                    const int64_t iw = ow * p.sw - p.pw + fw * p.dw;
                    if (iw < 0 || iw >= p.W)
                        continue;
                    acc += Elem<T>::ld(ip + ih * p.in_s[2] + iw * p.in_s[3]) * Elem<T>::ld(wp + fh * p.w_s[2] + fw * p.w_s[3]);
                }
            }
        }
        if (p.bias)
            acc += Elem<T>::ld(p.bias + oc * p.bias_s);
        Elem<T>::st(p.out + i * p.esize, acc);
    }
}

// weight layout (C_in, C_out / groups, KH, KW)
template <typename T> __global__ void __launch_bounds__(256) conv_transpose_kernel(const __grid_constant__ ConvParams p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ow = i % p.OW, t1 = i / p.OW, oh = t1 % p.OH, t2 = t1 / p.OH, oc = t2 % p.Cout, n = t2 / p.Cout;
        const int64_t g = oc / p.cout_g, ocg = oc - g * p.cout_g;
        double acc = 0.0; // always holds a value of type T
        for (int64_t ic = g * p.cin_g; ic < (g + 1) * p.cin_g; ++ic) {
            const char *ip = p.in + n * p.in_s[0] + ic * p.in_s[1];
            const char *wp = p.w + ic * p.w_s[0] + ocg * p.w_s[1];
            for (int64_t fh = p.KH - 1; fh >= 0; --fh) { // ih ascending
                const int64_t th = oh + p.ph - fh * p.dh;
                if (th < 0 || th % p.sh != 0 || th / p.sh >= p.H)
                    continue;
                const int64_t ih = th / p.sh;
                for (int64_t fw = p.KW - 1; fw >= 0; --fw) { // iw ascending
                    const int64_t tw = ow + p.pw - fw * p.dw;
                    if (tw < 0 || tw % p.sw != 0 || tw / p.sw >= p.W)
                        continue;
                    const int64_t iw = tw / p.sw;
                    const double prod = Elem<T>::round(Elem<T>::ld(ip + ih * p.in_s[2] + iw * p.in_s[3]) *
                                                       Elem<T>::ld(wp + fh * p.w_s[2] + fw * p.w_s[3]));
                    acc = Elem<T>::round(acc + prod);
                }
            }
        }
        if (p.bias)
            acc = Elem<T>::round(acc + Elem<T>::ld(p.bias + oc * p.bias_s));
        Elem<T>::st(p.out + i * p.esize, acc);
    }
}

// ---- interpolate ------------------------------------------------------------------------------------------------------
struct InterpParams {
    const char *in;
    char *out;
    int64_t in_s[4]; // byte strides (N, C, H, W); 1-D inputs use H = 1
    int64_t NC, C, H, W, OH, OW;
    int32_t mode, align, esize, one_d;
    int64_t total;
};

// The CPU build has no FMA contraction (plain x86-64 code generation): every product and sum below is rounded on its own,
// which the _rn intrinsics guarantee here (nvcc would otherwise fuse them and differ in the last bit).
#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
__device__ __forceinline__ float src_coord(int64_t o, int64_t in, int64_t out, int align) {
    if (align && out > 1)
        return __fdiv_rn(FM((float)o, (float)(in - 1)), (float)(out - 1));
    return FS(__fdiv_rn(FM(FA((float)o, 0.5f), (float)in), (float)out), 0.5f);
}
__device__ __forceinline__ float cubic_w(float x) { // a = -0.75 (operations.cpp:3915-3928), constants folded as the compiler folds them
    x = fabsf(x);
    if (x <= 1.0f)
        return FA(FS(FM(FM(FM(1.25f, x), x), x), FM(FM(2.25f, x), x)), 1.0f);
    if (x < 2.0f)
        return FS(FA(FS(FM(FM(FM(-0.75f, x), x), x), FM(FM(-3.75f, x), x)), FM(-6.0f, x)), -3.0f);
    return 0.0f;
}

template <typename T> __global__ void __launch_bounds__(256) interpolate_kernel(const __grid_constant__ InterpParams p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ow = i % p.OW, t1 = i / p.OW, oh = t1 % p.OH, nc = t1 / p.OH;
        const int64_t n = nc / p.C, c = nc - n * p.C;
        const char *ip = p.in + n * p.in_s[0] + c * p.in_s[1];
        char *op = p.out + i * p.esize;
        auto at = [&](int64_t h, int64_t w) { return ip + h * p.in_s[2] + w * p.in_s[3]; };
        if (p.mode == 0) { // nearest: floor(o * in / out) in fp32, clamped
            int64_t sh = 0, sw = (int64_t)__fdiv_rn(FM((float)ow, (float)p.W), (float)p.OW);
            if (!p.one_d)
                sh = (int64_t)__fdiv_rn(FM((float)oh, (float)p.H), (float)p.OH);
            sh = sh >= p.H ? p.H - 1 : sh;
            sw = sw >= p.W ? p.W - 1 : sw;
            // a copy: move the element's bytes
            if (p.esize == 8)
                *reinterpret_cast<uint64_t *>(op) = *reinterpret_cast<const uint64_t *>(at(sh, sw));
            else if (p.esize == 4)
                *reinterpret_cast<uint32_t *>(op) = *reinterpret_cast<const uint32_t *>(at(sh, sw));
            else
                *reinterpret_cast<uint16_t *>(op) = *reinterpret_cast<const uint16_t *>(at(sh, sw));
            continue;
        }
        float fw = src_coord(ow, p.W, p.OW, p.align);
        float fh = p.one_d ? 0.0f : src_coord(oh, p.H, p.OH, p.align);
        if (p.mode == 1 || p.one_d) { // linear (1-D) / bilinear
            fw = fmaxf(0.0f, fminf(fw, (float)(p.W - 1)));
            fh = fmaxf(0.0f, fminf(fh, (float)(p.H - 1)));
            const int64_t w0 = (int64_t)fw, h0 = (int64_t)fh;
            const int64_t w1 = min(w0 + 1, p.W - 1), h1 = min(h0 + 1, p.H - 1);
            const float ww = FS(fw, (float)w0), wh = FS(fh, (float)h0);
            float val;
            if (p.one_d) {
                val = FA(FM(FS(1.0f, ww), Elem<T>::ldf(at(0, w0))), FM(ww, Elem<T>::ldf(at(0, w1))));
            } else {
                const float v00 = Elem<T>::ldf(at(h0, w0)), v01 = Elem<T>::ldf(at(h0, w1));
                const float v10 = Elem<T>::ldf(at(h1, w0)), v11 = Elem<T>::ldf(at(h1, w1));
                const float ah = FS(1.0f, wh), aw = FS(1.0f, ww);
                val = FA(FA(FA(FM(FM(ah, aw), v00), FM(FM(ah, ww), v01)), FM(FM(wh, aw), v10)), FM(FM(wh, ww), v11));
            }
            Elem<T>::stf(op, val);
            continue;
        }
        // bicubic (a = -0.75), border replicated
        const int hf = (int)floorf(fh), wf = (int)floorf(fw);
        float val = 0.0f;
#pragma unroll
        for (int dh = -1; dh <= 2; ++dh) {
            const int sh = max(0, min(hf + dh, (int)p.H - 1));
            const float wh = cubic_w(FS(fh, (float)(hf + dh)));
#pragma unroll
            for (int dw = -1; dw <= 2; ++dw) {
                const int sw = max(0, min(wf + dw, (int)p.W - 1));
                const float ww = cubic_w(FS(fw, (float)(wf + dw)));
                val = FA(val, FM(FM(wh, ww), Elem<T>::ldf(at(sh, sw))));
            }
        }
        Elem<T>::stf(op, val);
    }
}

unsigned grid_for(int64_t total, const DeviceCtx *ctx) {
    const int64_t blocks = (total + 255) / 256;
    return (unsigned)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 16));
}

#define AXB_DISPATCH_FLOAT(dt, KERNEL, grid, s, params)                                                                \
    switch (dt) {                                                                                                      \
    case AXB_F32: KERNEL<float><<<grid, 256, 0, s>>>(params); break;                                                   \
    case AXB_F64: KERNEL<double><<<grid, 256, 0, s>>>(params); break;                                                  \
    case AXB_BF16: KERNEL<BF16><<<grid, 256, 0, s>>>(params); break;                                                   \
    default: KERNEL<F16><<<grid, 256, 0, s>>>(params); break;                                                          \
    }

// Transposed convolution on the tensor cores. out[n, co, ih*sh - ph + kh*dh, iw*sw - pw + kw*dw] += x[n, ci, ih, iw] * w[ci, co, kh, kw]
// (ops::conv_transpose1d / 2d, src/tensor/operations.cpp:3425-3690) is, for the outputs of ONE phase (oh % sh, ow % sw) = (rh, rw),
// an ordinary stride-1 convolution of the un-dilated input: with oh = oh' * sh + rh only the taps with (rh + ph - kh*dh) % sh == 0
// contribute, they form an arithmetic progression kh = kh0 + i * step, and tap i reads input row oh' + q0 - i * (step * dh / sh).
// So every phase is one call of the implicit-GEMM convolution (gemm::conv_tc) with
//   * the weights seen through a strided VIEW (C_out / C_in strides swapped, taps kh0 :: step) — no re-layout pass of its own,
//   * a negative "dilation" and "padding" -q0 (conv_tc only turns them into per-tap TMA coordinates; rows / columns outside the
//     input are zero-filled, which is exactly the scatter's bounds test),
//   * the result written into the interleaved view out[:, :, rh :: sh, rw :: sw] (bias added once per element in the epilogue).
// Phases without taps (kernel smaller than the stride) and small problems stay on the gather kernel. Numerics as for conv_tc:
// fp32 through the exact three-piece bf16 split (fp32-FMA accuracy, tolerance-based parity, not the fp64 products of the loop).
// Groups: each group is the same computation on channel slices (input channels g*cin_g.., weight rows g*cin_g.., result
// channels g*cout_g..), taken as views.
axb_status conv_transpose_tc(const ConvParams &p, int dtype, int groups, DeviceCtx *ctx, cudaStream_t s, bool *done) {
    *done = false;
    static const bool off = getenv("AXB_CONVT_NO_TC") != nullptr; // A/B switch
    if (off || !(dtype == AXB_F32 || dtype == AXB_BF16 || dtype == AXB_F16))
        return AXB_OK;
    const double flops = 2.0 * (double)p.N * (double)p.cin_g * (double)p.H * (double)p.W * (double)p.cout_g * (double)(p.KH * p.KW);
    if (flops < 2.0e8 || p.cout_g < 32 || p.cin_g < 8 || p.OW < 8 * p.sw || p.sh * p.sw > 16 || groups > 64)
        return AXB_OK; // (sized per group: every group pays its own preparation passes)
    struct Axis {
        int64_t k0, step, count, q0, dil; // first tap, tap spacing, taps, input offset of the first tap, offset change per tap
    };
    auto plan = [](int64_t r, int64_t stride, int64_t pad, int64_t dil, int64_t K, Axis *a) {
        a->count = 0;
        int64_t prev = -1;
        for (int64_t k = 0; k < K; ++k) {
            const int64_t t = r + pad - k * dil;
            if (((t % stride) + stride) % stride != 0)
                continue;
            if (a->count == 0) {
                a->k0 = k;
                a->q0 = t / stride; // exact
                a->step = 1;
                a->dil = 0;
            } else if (a->count == 1) {
                a->step = k - prev;
                a->dil = t / stride - a->q0;
            }
            prev = k;
            a->count++;
        }
    };
    const int esz = (int)dtype_size(dtype);
    // every phase must have taps and rows, otherwise the gather kernel does the whole tensor
    for (int64_t rh = 0; rh < p.sh; ++rh)
        for (int64_t rw = 0; rw < p.sw; ++rw) {
            Axis ah, aw;
            plan(rh, p.sh, p.ph, p.dh, p.KH, &ah);
            plan(rw, p.sw, p.pw, p.dw, p.KW, &aw);
            if (ah.count == 0 || aw.count == 0 || rh >= p.OH || rw >= p.OW)
                return AXB_OK;
        }
    bool all = true;
    for (int g = 0; g < groups && all; ++g) {
    void *x_prepared = nullptr; // channels-last (split) form of the input: made by the first phase, read by all of them
    for (int64_t rh = 0; rh < p.sh && all; ++rh)
        for (int64_t rw = 0; rw < p.sw && all; ++rw) {
            Axis ah, aw;
            plan(rh, p.sh, p.ph, p.dh, p.KH, &ah);
            plan(rw, p.sw, p.pw, p.dw, p.KW, &aw);
            gemm::ConvDesc cd;
            memset(&cd, 0, sizeof cd);
            cd.x = p.in + g * p.cin_g * p.in_s[1];
            cd.w = p.w + g * p.cin_g * p.w_s[0] + ah.k0 * p.w_s[2] + aw.k0 * p.w_s[3];
            cd.bias = p.bias ? p.bias + g * p.cout_g * p.bias_s : nullptr;
            cd.out = p.out + (g * p.cout_g * p.OH * p.OW + rh * p.OW + rw) * esz;
            cd.dtype = dtype; cd.bias_dtype = dtype; cd.act = 0;
            cd.N = p.N; cd.Cin = p.cin_g; cd.H = p.H; cd.W = p.W; cd.Cout = p.cout_g;
            cd.KH = ah.count; cd.KW = aw.count;
            cd.OH = (p.OH - rh + p.sh - 1) / p.sh; cd.OW = (p.OW - rw + p.sw - 1) / p.sw;
            for (int i = 0; i < 4; ++i)
                cd.x_s[i] = p.in_s[i];
            cd.w_s[0] = p.w_s[1]; cd.w_s[1] = p.w_s[0]; // (C_in, C_out, ..) read as (C_out, C_in, ..)
            cd.w_s[2] = ah.step * p.w_s[2]; cd.w_s[3] = aw.step * p.w_s[3];
            cd.bias_s = p.bias_s;
            cd.sh = 1; cd.sw = 1;
            cd.ph = (int32_t)-ah.q0; cd.pw = (int32_t)-aw.q0;
            cd.dh = (int32_t)ah.dil; cd.dw = (int32_t)aw.dil;
            cd.o_s[0] = p.Cout * p.OH * p.OW * esz; cd.o_s[1] = p.OH * p.OW * esz; cd.o_s[2] = p.sh * p.OW * esz; cd.o_s[3] = p.sw * esz;
            cd.force = 1;
            cd.x_prepared = &x_prepared;
            bool one = false;
            const axb_status st = gemm::conv_tc(cd, ctx, s, &one);
            if (st != AXB_OK) {
                if (x_prepared)
                    axb_free(x_prepared, s);
                return st;
            }
            all = all && one;
        }
    if (x_prepared)
        axb_free(x_prepared, s); // stream-ordered: after the last phase's GEMM
    }
    *done = all; // a phase the tensor-core path declined: the gather kernel rewrites the whole result
    return AXB_OK;
}

// shared argument handling of conv2d / conv_transpose: 3-D (N, C, L) tensors are the 1-D case (H = KH = 1)
axb_status run_conv(bool transposed, const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, int spatial_dims,
                    const int32_t *stride, const int32_t *padding, const int32_t *dilation, int groups, const axb_tensor *out,
                    axb_stream_t stream, int activation = 0) {
    const char *what = transposed ? "conv_transpose" : "conv2d";
    AXB_TRY(validate_tensor(in, what));
    AXB_TRY(validate_tensor(weight, what));
    AXB_TRY(validate_tensor(out, what));
    AXB_REQUIRE(spatial_dims == 1 || spatial_dims == 2, AXB_ERR_INVALID, "%s: 1 or 2 spatial dims", what);
    const int nd = spatial_dims + 2;
    AXB_REQUIRE(in->ndim == nd && weight->ndim == nd && out->ndim == nd, AXB_ERR_SHAPE, "%s: expected %d-d input, weight and output",
                what, nd);
    AXB_REQUIRE(is_float_dtype(in->dtype) && weight->dtype == in->dtype && out->dtype == in->dtype, AXB_ERR_UNSUPPORTED,
                "%s: input, weight and output must share one floating dtype", what);
    AXB_REQUIRE(is_contiguous(*out), AXB_ERR_UNSUPPORTED, "%s: output must be contiguous", what);
    AXB_REQUIRE(groups > 0, AXB_ERR_INVALID, "%s: groups must be > 0", what);
    ConvParams p;
    memset(&p, 0, sizeof p);
    const int o = spatial_dims == 1 ? 1 : 0; // where the (possibly missing) H axis sits
    p.in = tensor_ptr(*in); p.w = tensor_ptr(*weight); p.out = tensor_ptr(*out);
    p.N = in->shape[0]; p.Cin = in->shape[1];
    p.H = o ? 1 : in->shape[2]; p.W = in->shape[nd - 1];
    p.KH = o ? 1 : weight->shape[2]; p.KW = weight->shape[nd - 1];
    p.Cout = out->shape[1]; p.OH = o ? 1 : out->shape[2]; p.OW = out->shape[nd - 1];
    p.in_s[0] = in->strides[0]; p.in_s[1] = in->strides[1];
    p.in_s[2] = o ? 0 : in->strides[2]; p.in_s[3] = in->strides[nd - 1];
    p.w_s[0] = weight->strides[0]; p.w_s[1] = weight->strides[1];
    p.w_s[2] = o ? 0 : weight->strides[2]; p.w_s[3] = weight->strides[nd - 1];
    p.sh = o ? 1 : stride[0]; p.sw = stride[spatial_dims - 1];
    p.ph = o ? 0 : padding[0]; p.pw = padding[spatial_dims - 1];
    p.dh = o ? 1 : dilation[0]; p.dw = dilation[spatial_dims - 1];
    AXB_REQUIRE(p.sh > 0 && p.sw > 0 && p.dh > 0 && p.dw > 0 && p.ph >= 0 && p.pw >= 0, AXB_ERR_INVALID,
                "%s: stride / dilation must be > 0 and padding >= 0", what);
    AXB_REQUIRE(out->shape[0] == p.N, AXB_ERR_SHAPE, "%s: batch mismatch", what);
    AXB_REQUIRE(p.Cin % groups == 0 && p.Cout % groups == 0, AXB_ERR_SHAPE, "%s: channels not divisible by groups", what);
    p.cin_g = p.Cin / groups; p.cout_g = p.Cout / groups;
    if (!transposed) // weight (C_out, C_in / groups, ..)
        AXB_REQUIRE(weight->shape[0] == p.Cout && weight->shape[1] == p.cin_g, AXB_ERR_SHAPE, "conv2d: weight shape does not match the channels");
    else             // weight (C_in, C_out / groups, ..)
        AXB_REQUIRE(weight->shape[0] == p.Cin && weight->shape[1] == p.cout_g, AXB_ERR_SHAPE,
                    "conv_transpose: weight shape does not match the channels");
    if (bias) {
        AXB_TRY(validate_tensor(bias, what));
        AXB_REQUIRE(bias->dtype == in->dtype && bias->ndim == 1 && bias->shape[0] == p.Cout, AXB_ERR_SHAPE, "%s: bias must be {C_out} of the input dtype", what);
        p.bias = tensor_ptr(*bias);
        p.bias_s = bias->strides[0];
    }
    p.esize = (int32_t)dtype_size(in->dtype);
    p.total = numel(*out);
    if (p.total == 0)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    if (!transposed && groups == 1) {
        // large problems: implicit GEMM on the tensor cores (gemm_tcgen05.cu conv_tc) — fp32 through the exact three-piece
        // bf16 split (as accurate as an fp32 FMA chain, NOT bit-identical to the fp64 accumulation of the gather kernel)
        gemm::ConvDesc cd;
        memset(&cd, 0, sizeof cd);
        cd.x = p.in; cd.w = p.w; cd.bias = p.bias; cd.out = p.out;
        cd.dtype = in->dtype; cd.bias_dtype = in->dtype; cd.act = activation;
        cd.N = p.N; cd.Cin = p.Cin; cd.H = p.H; cd.W = p.W; cd.Cout = p.Cout; cd.KH = p.KH; cd.KW = p.KW; cd.OH = p.OH; cd.OW = p.OW;
        for (int i = 0; i < 4; ++i) {
            cd.x_s[i] = p.in_s[i];
            cd.w_s[i] = p.w_s[i];
        }
        cd.bias_s = p.bias_s;
        cd.sh = p.sh; cd.sw = p.sw; cd.ph = p.ph; cd.pw = p.pw; cd.dh = p.dh; cd.dw = p.dw;
        bool done = false;
        AXB_TRY(gemm::conv_tc(cd, ctx, s, &done));
        if (done)
            return AXB_OK;
    }
    if (!transposed && groups > 1 && groups <= 64 && p.cin_g >= 8 && p.cout_g >= 32) {
        // grouped convolution: every group is a convolution of a channel slice of the input into a channel slice of the
        // result — views of the same tensors, so the tensor-core path runs once per group (each call sizes itself; if one
        // declines, the gather kernel below rewrites the whole result)
        bool all = true;
        for (int g = 0; g < groups && all; ++g) {
            gemm::ConvDesc cd;
            memset(&cd, 0, sizeof cd);
            cd.x = p.in + g * p.cin_g * p.in_s[1];
            cd.w = p.w + g * p.cout_g * p.w_s[0];
            cd.bias = p.bias ? p.bias + g * p.cout_g * p.bias_s : nullptr;
            cd.out = p.out + g * p.cout_g * p.OH * p.OW * p.esize;
            cd.dtype = in->dtype; cd.bias_dtype = in->dtype; cd.act = activation;
            cd.N = p.N; cd.Cin = p.cin_g; cd.H = p.H; cd.W = p.W; cd.Cout = p.cout_g; cd.KH = p.KH; cd.KW = p.KW; cd.OH = p.OH; cd.OW = p.OW;
            for (int i = 0; i < 4; ++i) {
                cd.x_s[i] = p.in_s[i];
                cd.w_s[i] = p.w_s[i];
            }
            cd.bias_s = p.bias_s;
            cd.sh = p.sh; cd.sw = p.sw; cd.ph = p.ph; cd.pw = p.pw; cd.dh = p.dh; cd.dw = p.dw;
            cd.o_s[0] = p.Cout * p.OH * p.OW * p.esize; cd.o_s[1] = p.OH * p.OW * p.esize; cd.o_s[2] = p.OW * p.esize; cd.o_s[3] = p.esize;
            bool one = false;
            AXB_TRY(gemm::conv_tc(cd, ctx, s, &one));
            all = all && one;
        }
        if (all)
            return AXB_OK;
    }
    if (transposed) {
        bool done = false;
        AXB_TRY(conv_transpose_tc(p, in->dtype, groups, ctx, s, &done));
        if (done)
            return AXB_OK;
    }
    gemm::last_path() = 0;
    const unsigned grid = grid_for(p.total, ctx);
    if (transposed) {
        AXB_DISPATCH_FLOAT(in->dtype, conv_transpose_kernel, grid, s, p)
    } else {
        AXB_DISPATCH_FLOAT(in->dtype, conv2d_kernel, grid, s, p)
    }
    AXB_CHECK_LAUNCH();
    if (activation) // the gather kernel has no epilogue: the activation runs in place on the result (same rounding order as
        AXB_TRY(axb_unary(activation, out, out, activation == AXB_OP_LEAKY_RELU ? 0.01 : 0.0, stream)); // conv followed by the op)
    return AXB_OK;
}

} // namespace
} // namespace axb

using namespace axb;

extern "C" {

axb_status axb_conv2d(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, const int32_t stride[2],
                      const int32_t padding[2], const int32_t dilation[2], int groups, const axb_tensor *out, axb_stream_t stream) {
    return run_conv(false, in, weight, bias, 2, stride, padding, dilation, groups, out, stream);
}

axb_status axb_conv(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, int spatial_dims, const int32_t *stride,
                    const int32_t *padding, const int32_t *dilation, int groups, int activation, const axb_tensor *out,
                    axb_stream_t stream) {
    AXB_REQUIRE(stride && padding && dilation, AXB_ERR_INVALID, "axb_conv: null stride / padding / dilation");
    return run_conv(false, in, weight, bias, spatial_dims, stride, padding, dilation, groups, out, stream, activation);
}

axb_status axb_conv_transpose(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, int spatial_dims,
                              const int32_t *stride, const int32_t *padding, const int32_t *dilation, int groups,
                              const axb_tensor *out, axb_stream_t stream) {
    return run_conv(true, in, weight, bias, spatial_dims, stride, padding, dilation, groups, out, stream);
}

axb_status axb_interpolate(const axb_tensor *in, int mode, int align_corners, const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "interpolate"));
    AXB_TRY(validate_tensor(out, "interpolate"));
    AXB_REQUIRE(in->ndim == 3 || in->ndim == 4, AXB_ERR_SHAPE, "interpolate: only 1D and 2D spatial interpolation supported");
    AXB_REQUIRE(out->ndim == in->ndim && out->shape[0] == in->shape[0] && out->shape[1] == in->shape[1], AXB_ERR_SHAPE,
                "interpolate: output must be (N, C, ..target size..)");
    AXB_REQUIRE(is_float_dtype(in->dtype) && out->dtype == in->dtype, AXB_ERR_UNSUPPORTED, "interpolate: floating dtypes only");
    AXB_REQUIRE(mode >= 0 && mode <= 2, AXB_ERR_INVALID, "interpolate: mode 0 (nearest), 1 (linear / bilinear) or 2 (bicubic)");
    AXB_REQUIRE(is_contiguous(*out), AXB_ERR_UNSUPPORTED, "interpolate: output must be contiguous");
    InterpParams p;
    memset(&p, 0, sizeof p);
    const int nd = in->ndim;
    p.one_d = nd == 3;
    p.in = tensor_ptr(*in); p.out = tensor_ptr(*out);
    p.C = in->shape[1]; p.NC = in->shape[0] * in->shape[1];
    p.H = p.one_d ? 1 : in->shape[2]; p.W = in->shape[nd - 1];
    p.OH = p.one_d ? 1 : out->shape[2]; p.OW = out->shape[nd - 1];
    p.in_s[0] = in->strides[0]; p.in_s[1] = in->strides[1];
    p.in_s[2] = p.one_d ? 0 : in->strides[2]; p.in_s[3] = in->strides[nd - 1];
    p.mode = mode; p.align = align_corners ? 1 : 0;
    p.esize = (int32_t)dtype_size(in->dtype);
    p.total = numel(*out);
    if (p.total == 0)
        return AXB_OK;
    AXB_REQUIRE(p.H > 0 && p.W > 0, AXB_ERR_SHAPE, "interpolate: empty input");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    const unsigned grid = grid_for(p.total, ctx);
    AXB_DISPATCH_FLOAT(in->dtype, interpolate_kernel, grid, s, p)
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

} // extern "C"
