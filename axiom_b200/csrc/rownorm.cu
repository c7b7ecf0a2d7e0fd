// rownorm.cu — kernel family 2b: softmax / log_softmax / layer_norm / rms_norm along one axis.
//
// Replaces CPUSoftmaxOperation (src/backends/cpu/cpu_operations.cpp:3067-3241) and ops::layer_norm / rms_norm
// (src/tensor/operations.cpp:1775-1858; the Metal backend's fused formula is mpsgraph_graph_compiler.mm:303-330 and its
// hand-written layer_norm_fwd kernel src/backends/metal/kernels.metal:167-411).
//
// HBM-bound: each row is read ONCE, kept in REGISTERS (fp32, up to 32 values per thread), reduced with warp
// shuffles (+ one shared-memory hop for multi-warp rows) and written once: algorithmic traffic = in + out
// (SURVEY.md §8d C3). Three resident shapes cover rows up to 32 768 elements: one warp per row (<= 1024), 256 threads
// per row (<= 8192), 1024 threads per row (<= 32 768: 128 KiB of fp32 = half the SM's register file). Longer rows
// stream from global memory in three passes. Contiguous rows use 128-bit (fp32) / 64-bit (half) streaming accesses.
//
// Numerics mirror the reference op sequence, every step rounded to fp32 (no FMA contraction):
//   softmax    m = max, e = exp(x - m), s = sum(e), y = e / s           (log: y = x - m - log(s))   [:3123-3162]
//   layer_norm mean = sum(x)/n; xc = x - mean; var = sum(xc*xc)/n; y = xc / sqrt(var + eps) * w + b  [:1803-1817]
//   rms_norm   ms = sum(x*x)/n; y = x / sqrt(ms + eps) * w                                           [:1841-1857]
// Half inputs: the CPU path rounds mean (layer_norm) and x*x, ms (rms_norm) to the half type because those steps run
// in the input dtype (bfloat16.hpp:55-66, cpu_operations.cpp:1613-1614); mirrored here. Sums are trees, not the
// reference's serial / lane-strided order: results agree to rounding (tests use the north_star's 1e-5 bound).

#include "rownorm_common.cuh"
#include "rownorm_wide.cuh"

namespace axb {
namespace rown {

// THREADS threads per row (32: one warp per row, 256/THREADS... rows per CTA), VPT resident values per thread.
// RESIDENT = false: rows longer than THREADS*VPT re-read global memory in every pass.
template <typename In, typename Out, int THREADS, int VPT, int KIND, bool RESIDENT>
__global__ void __launch_bounds__(THREADS == 32 ? 256 : THREADS) rownorm_kernel(const __grid_constant__ RowParams p) {
    constexpr int CTA = THREADS == 32 ? 256 : THREADS;
    constexpr int ROWS = CTA / THREADS;
    constexpr int NCH = VPT / 4; // 4-element chunks per thread
    __shared__ float smem[32];
    const int tid = threadIdx.x % THREADS;
    const int n = p.n;
    const float inv_n = 0.0f; (void)inv_n;
    for (int64_t row0 = (int64_t)blockIdx.x * ROWS; row0 < p.rows; row0 += (int64_t)gridDim.x * ROWS) {
        const int64_t row = row0 + threadIdx.x / THREADS;
        const bool row_ok = row < p.rows; // only differs inside a CTA when THREADS == 32 (no __syncthreads there)
        int64_t in_off = 0, out_off = 0;
        if (row_ok)
            row_offsets(p, row, &in_off, &out_off);
        const char *src = p.in + in_off;
        char *dst = p.out + out_off;
        float x[RESIDENT ? VPT : 4];

        // element access helpers -------------------------------------------------------------------
        auto load_chunk = [&](int e0, float (&v)[4]) { // elements e0..e0+3 (e0 % 4 == 0)
            if (p.vec_ok && e0 + 4 <= n) {
                ld4<In>(src + (int64_t)e0 * sizeof(In), v);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    v[k] = (e0 + k < n) ? ld_f<In>(src + (int64_t)(e0 + k) * p.in_es) : 0.0f;
            }
        };
        auto store_chunk = [&](int e0, const float (&v)[4]) {
            if (p.vec_ok && e0 + 4 <= n) {
                st4<Out>(dst + (int64_t)e0 * sizeof(Out), v);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (e0 + k < n)
                        st_f<Out>(dst + (int64_t)(e0 + k) * p.out_es, v[k]);
            }
        };
        // visits every element of the row owned by this thread: f(e, value&) — value is the resident register or a
        // fresh load; WRITE stores f's updated value to `dst`
        auto sweep = [&](auto f, bool reload, bool write) {
            if (RESIDENT) {
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int e0 = (j * THREADS + tid) * 4;
                    if (e0 < n && row_ok) {
                        float v[4];
                        if (reload) {
                            load_chunk(e0, v);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                x[(RESIDENT ? j * 4 : 0) + k] = v[k];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            v[k] = x[(RESIDENT ? j * 4 : 0) + k];
                            if (e0 + k < n)
                                f(e0 + k, v[k]);
                            x[(RESIDENT ? j * 4 : 0) + k] = v[k];
                        }
                        if (write)
                            store_chunk(e0, v);
                    }
                }
            } else {
                for (int e0 = tid * 4; e0 < n && row_ok; e0 += THREADS * 4) {
                    float v[4];
                    load_chunk(e0, v);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (e0 + k < n)
                            f(e0 + k, v[k]);
                    if (write)
                        store_chunk(e0, v);
                }
            }
        };

        if (KIND == K_SOFTMAX || KIND == K_LOG_SOFTMAX) {
            float m = -INFINITY;
            sweep([&](int, float &v) { m = fmaxf(m, v); }, true, false);
            m = row_allreduce<THREADS, FMax>(m, smem);
            float s = 0.0f;
            if (KIND == K_SOFTMAX) {
                // one IEEE reciprocal per row instead of a division per element, and the 6-instruction exp of the wide
                // kernel (rownorm_wide.cuh w_exp_neg): both inside the 1e-5 contract the tests state
                if (RESIDENT) { // e = exp(x - m) kept in registers, y = e * (1 / s)
                    sweep([&](int, float &v) { v = w_exp_neg(__fsub_rn(v, m)); s += v; }, false, false);
                    s = row_allreduce<THREADS, FAdd>(s, smem);
                    const float inv = __fdiv_rn(1.0f, s);
                    sweep([&](int, float &v) { v = __fmul_rn(v, inv); }, false, true);
                } else {
                    sweep([&](int, float &v) { s += w_exp_neg(__fsub_rn(v, m)); }, true, false);
                    s = row_allreduce<THREADS, FAdd>(s, smem);
                    const float inv = __fdiv_rn(1.0f, s);
                    sweep([&](int, float &v) { v = __fmul_rn(w_exp_neg(__fsub_rn(v, m)), inv); }, true, true);
                }
            } else {
                sweep([&](int, float &v) { s += w_exp_neg(__fsub_rn(v, m)); }, !RESIDENT, false);
                s = row_allreduce<THREADS, FAdd>(s, smem);
                const float ls = logf(s);
                sweep([&](int, float &v) { v = __fsub_rn(__fsub_rn(v, m), ls); }, !RESIDENT, true);
            }
        } else if (KIND == K_LAYER_NORM) {
            float s = 0.0f;
            sweep([&](int, float &v) { s += v; }, true, false);
            s = row_allreduce<THREADS, FAdd>(s, smem);
            const float mean = round_to<In>(__fdiv_rn(s, (float)n));
            float q = 0.0f;
            sweep([&](int, float &v) { v = __fsub_rn(v, mean); q += __fmul_rn(v, v); }, !RESIDENT, false);
            q = row_allreduce<THREADS, FAdd>(q, smem);
            const float rsd = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(q, (float)n), p.eps)));
            // contiguous Float32 parameters (nn::LayerNorm): plain indexed loads instead of the dtype switch + 64-bit
            // stride arithmetic per element
            const bool w_fast = p.w && p.w_dtype == AXB_F32 && p.w_es == 4, b_fast = p.b && p.b_dtype == AXB_F32 && p.b_es == 4;
            const float *wf = reinterpret_cast<const float *>(p.w), *bf = reinterpret_cast<const float *>(p.b);
            if (RESIDENT && p.vec_ok && (w_fast || !p.w) && (b_fast || !p.b) && (n & 3) == 0 &&
                ((reinterpret_cast<uintptr_t>(p.w) | reinterpret_cast<uintptr_t>(p.b)) & 15) == 0) {
                // final pass with 128-bit parameter loads: the per-element scalar loads of weight and bias were the
                // largest L1 client of this kernel on transformer-width rows (2 x 32-bit loads per output element)
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int e0 = (j * THREADS + tid) * 4;
                    if (e0 < n && row_ok) {
                        float y[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            y[k] = __fmul_rn(x[(RESIDENT ? j * 4 : 0) + k], rsd);
                        if (p.w) {
                            const float4 wv = __ldg(reinterpret_cast<const float4 *>(wf + e0));
                            y[0] = __fmul_rn(y[0], wv.x); y[1] = __fmul_rn(y[1], wv.y); y[2] = __fmul_rn(y[2], wv.z); y[3] = __fmul_rn(y[3], wv.w);
                        }
                        if (p.b) {
                            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bf + e0));
                            y[0] = __fadd_rn(y[0], bv.x); y[1] = __fadd_rn(y[1], bv.y); y[2] = __fadd_rn(y[2], bv.z); y[3] = __fadd_rn(y[3], bv.w);
                        }
                        store_chunk(e0, y);
                    }
                }
            } else {
            sweep([&](int e, float &v) {
                float xc = RESIDENT ? v : __fsub_rn(v, mean);
                float y = __fmul_rn(xc, rsd);
                if (w_fast)
                    y = __fmul_rn(y, __ldg(wf + e));
                else if (p.w)
                    y = __fmul_rn(y, ld_dyn(p.w + (int64_t)e * p.w_es, p.w_dtype));
                if (b_fast)
                    y = __fadd_rn(y, __ldg(bf + e));
                else if (p.b)
                    y = __fadd_rn(y, ld_dyn(p.b + (int64_t)e * p.b_es, p.b_dtype));
                v = y;
            }, !RESIDENT, true);
            }
        } else { // K_RMS_NORM
            float q = 0.0f;
            sweep([&](int, float &v) { q += round_to<In>(__fmul_rn(v, v)); }, true, false);
            q = row_allreduce<THREADS, FAdd>(q, smem);
            const float ms = round_to<In>(__fdiv_rn(q, (float)n));
            const float rrms = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(ms, p.eps)));
            const bool w_fast = p.w && p.w_dtype == AXB_F32 && p.w_es == 4;
            const float *wf = reinterpret_cast<const float *>(p.w);
            sweep([&](int e, float &v) {
                float y = __fmul_rn(v, rrms);
                if (w_fast)
                    y = __fmul_rn(y, __ldg(wf + e));
                else if (p.w)
                    y = __fmul_rn(y, ld_dyn(p.w + (int64_t)e * p.w_es, p.w_dtype));
                v = y;
            }, !RESIDENT, true);
        }
    }
}


int &last_path() { // per thread: 0 generic (register-resident / streaming), 1 wide TMA-staged, 2 SM-pair cluster
    static thread_local int v = -1;
    return v;
}

namespace {

// Wide contiguous rows (>= 16 KiB, <= 32 values per lane of a 1024-thread CTA): persistent TMA-staged kernel.
template <typename In, typename Out, typename W, int KIND>
bool launch_wide_typed(RowParams p, DeviceCtx *ctx, cudaStream_t s, axb_status *st);

template <typename In, typename Out, int KIND> bool try_launch_wide(const RowParams &p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    // weight and bias must share one dtype class: the input's, or Float32 (what nn::LayerNorm holds for half inputs)
    const int wdt = p.w ? p.w_dtype : (p.b ? p.b_dtype : p.in_dtype);
    if (p.w && p.b && p.w_dtype != p.b_dtype)
        return false;
    if (wdt == p.in_dtype)
        return launch_wide_typed<In, Out, In, KIND>(p, ctx, s, st);
    if constexpr (KIND == K_LAYER_NORM || KIND == K_RMS_NORM) {
        if (wdt == AXB_F32)
            return launch_wide_typed<In, Out, float, KIND>(p, ctx, s, st);
    }
    return false;
}

template <typename In, typename Out, typename W, int KIND>
bool launch_wide_typed(RowParams p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    constexpr int E = 16 / (int)sizeof(In);
    const int64_t isz = sizeof(In), osz = sizeof(Out);
    if (p.in_es != isz || p.out_es != osz || ((int64_t)p.n * isz) % 16 != 0)
        return false;
    const int64_t nv = (int64_t)p.n * isz / 16;
    // rows of 32 KiB .. 128 KiB; shorter rows run several register-resident CTAs per SM (generic kernel), which hides
    // the reduction round trips better than one persistent CTA (measured: {32768, 4096} fp32 softmax 4.96 vs 3.09 TB/s)
    if (nv < 2048 || nv > 1024 * (32 / E))
        return false;
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p.in) || !al16(p.out))
        return false;
    for (int d = 0; d < p.nk; ++d)
        if (p.in_kstride[d] % 16 != 0 || p.out_kstride[d] % 16 != 0)
            return false;
    auto wb_ok = [&](const char *q, int64_t es, int dt) { return !q || (al16(q) && es == (int64_t)dtype_size(dt)); };
    if (!wb_ok(p.w, p.w_es, p.w_dtype) || !wb_ok(p.b, p.b_es, p.b_dtype))
        return false;
    static const bool disabled = getenv("AXB_NO_WIDE_ROWS") != nullptr; // A/B switch for profiling
    if (disabled)
        return false;
    p.nv = (int)nv;
    p.vpw = (int)((nv + kWideWarps - 1) / kWideWarps);
    const int j_need = (p.vpw + 31) / 32;
    const size_t budget = ctx->smem_optin - 4096; // static: mbarriers + reduction scratch
    const size_t row_bytes = (size_t)nv * 16;
    constexpr int JMAX = 32 / E;
    // one row per pass. (A variant that covered R = 2..8 medium rows per pass with shared block reductions measured
    // SLOWER on B200 — {16384, 8192} fp32 softmax 4.96 vs 5.30 TB/s, {16384, 16384} 5.71 vs 6.23 — and was removed.)
    constexpr int R = 1;
    const size_t per_stage = row_bytes * R;
    p.stages = (int)std::min<size_t>(kWideMaxStages, std::max<size_t>(1, budget / per_stage));
    if (per_stage > budget)
        return false;
    p.stages = (int)std::min<size_t>(p.stages, std::max<size_t>(1, (64 * 1024 + per_stage - 1) / per_stage)); // ~64 KiB in flight
    if (const char *e = getenv("AXB_WIDE_STAGES")) // probe switch: any depth the shared memory holds
        p.stages = std::max(1, std::min<int>((int)std::min<size_t>(kWideMaxStages, budget / per_stage), atoi(e)));
    static const bool no_alt = getenv("AXB_WB_NO_ALTERNATE") != nullptr;
    p.wb_alternate = no_alt ? 0 : 1;
    const size_t dyn = (size_t)p.stages * per_stage;
    const int64_t n_groups = (p.rows + R - 1) / R;
    const unsigned grid = (unsigned)std::min<int64_t>(n_groups, ctx->sm_count);
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "wide row kernel: cannot reserve %zu bytes of shared memory: %s", dyn, cudaGetErrorString(e));
            return;
        }
        kern<<<grid, kWideThreads, dyn, s>>>(p);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return;
        }
        count_launch();
        *st = AXB_OK;
    };
    const bool exact = p.vpw == JMAX * 32 && nv == (int64_t)kWideWarps * p.vpw;
    auto go_dyn = [&](auto kern, size_t bytes) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "wide row kernel: cannot reserve %zu bytes of shared memory: %s", bytes, cudaGetErrorString(e));
            return;
        }
        kern<<<grid, kWideThreads, bytes, s>>>(p);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return;
        }
        count_launch();
        *st = AXB_OK;
    };
    if constexpr (KIND == K_SOFTMAX || KIND == K_LOG_SOFTMAX) {
        if constexpr (sizeof(In) == 4) {
            // full-width fp32 rows (128 KiB, one row fills the ring): every warp fetches its 4 KiB slice as four 1 KiB
            // pieces and re-arms each piece as soon as it has been read. Measured on B200 at {8192, 32768}: 0.889 of the
            // copy peak with one 4 KiB copy per warp, 0.87 with two pieces, 1.009 with four (same 128 KiB ring, results
            // bit-identical), 0.77 with eight (512-byte copies).
            int pieces = 4;
            if (const char *e = getenv("AXB_WIDE_PIECES"))
                pieces = atoi(e);
            if (exact && (pieces == 2 || pieces == 4 || pieces == 8)) {
                const size_t slot = row_bytes / (size_t)pieces;
                int slots = (int)std::min<size_t>(kWideMaxSlots, budget / slot);
                slots = std::min(slots, pieces); // one row in flight: deeper rings measured the same or worse
                if (const char *e = getenv("AXB_WIDE_STAGES"))
                    slots = std::max(1, std::min<int>((int)std::min<size_t>(kWideMaxSlots, budget / slot), atoi(e)));
                p.stages = slots;
                const size_t bytes = (size_t)slots * slot;
                if (pieces == 2)
                    go_dyn(rowwide_kernel<In, Out, W, KIND, JMAX, true, 2>, bytes);
                else if (pieces == 4)
                    go_dyn(rowwide_kernel<In, Out, W, KIND, JMAX, true, 4>, bytes);
                else
                    go_dyn(rowwide_kernel<In, Out, W, KIND, JMAX, true, 8>, bytes);
                return true;
            }
        }
    }
    if (exact)
        go(rowwide_kernel<In, Out, W, KIND, JMAX, true>);
    else if (j_need * 4 <= JMAX && JMAX >= 4)
        go(rowwide_kernel<In, Out, W, KIND, (JMAX >= 4 ? JMAX / 4 : 1), false>);
    else if (j_need * 2 <= JMAX)
        go(rowwide_kernel<In, Out, W, KIND, JMAX / 2, false>);
    else
        go(rowwide_kernel<In, Out, W, KIND, JMAX, false>);
    return true;
}

template <typename In, typename Out, int KIND> axb_status launch_cfg(const RowParams &p, DeviceCtx *ctx, cudaStream_t s) {
    const int n = p.n;
    const int64_t max_ctas = (int64_t)ctx->sm_count * 32;
    axb_status wst = AXB_OK;
    if (try_launch_wide<In, Out, KIND>(p, ctx, s, &wst)) {
        last_path() = 1;
        return wst;
    }
    last_path() = 0;
    if (n <= 256) {
        int64_t g = std::min<int64_t>((p.rows + 7) / 8, max_ctas);
        rownorm_kernel<In, Out, 32, 8, KIND, true><<<(unsigned)g, 256, 0, s>>>(p);
    } else if (n <= 512) { // e.g. whisper's 384-wide layer norms: one warp per row, 16 values per lane
        int64_t g = std::min<int64_t>((p.rows + 7) / 8, max_ctas);
        rownorm_kernel<In, Out, 32, 16, KIND, true><<<(unsigned)g, 256, 0, s>>>(p);
    } else if (n <= 1024 && !getenv("AXB_ROWS_OLD_MEDIUM")) { // four warps per row, 8 values per thread
        int64_t g = std::min<int64_t>(p.rows, max_ctas * 2);
        rownorm_kernel<In, Out, 128, 8, KIND, true><<<(unsigned)g, 128, 0, s>>>(p);
    } else if (n <= 1024) {
        int64_t g = std::min<int64_t>((p.rows + 7) / 8, max_ctas);
        rownorm_kernel<In, Out, 32, 32, KIND, true><<<(unsigned)g, 256, 0, s>>>(p);
    } else if (n <= 2048) { // e.g. attention rows of 1500: 8 values per thread keep the unrolled sweeps short
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 256, 8, KIND, true><<<(unsigned)g, 256, 0, s>>>(p);
    } else if (n <= 4096 && !getenv("AXB_ROWS_OLD_MEDIUM")) { // 8 values per thread keep several CTAs resident per SM
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 512, 8, KIND, true><<<(unsigned)g, 512, 0, s>>>(p);
    } else if (n <= 8192 && !getenv("AXB_ROWS_OLD_MEDIUM")) {
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 1024, 8, KIND, true><<<(unsigned)g, 1024, 0, s>>>(p);
    } else if (n <= 8192) {
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 256, 32, KIND, true><<<(unsigned)g, 256, 0, s>>>(p);
    } else if (n <= 32768) {
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 1024, 32, KIND, true><<<(unsigned)g, 1024, 0, s>>>(p);
    } else {
        int64_t g = std::min<int64_t>(p.rows, max_ctas);
        rownorm_kernel<In, Out, 1024, 4, KIND, false><<<(unsigned)g, 1024, 0, s>>>(p);
    }
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

template <int KIND> axb_status dispatch(const RowParams &p, DeviceCtx *ctx, cudaStream_t s) {
    const int i = p.in_dtype, o = p.out_dtype;
    if (i == AXB_F32 && o == AXB_F32)
        return launch_cfg<float, float, KIND>(p, ctx, s);
    if (i == AXB_BF16 && o == AXB_BF16)
        return launch_cfg<bf16_t, bf16_t, KIND>(p, ctx, s);
    if (i == AXB_F16 && o == AXB_F16)
        return launch_cfg<f16_t, f16_t, KIND>(p, ctx, s);
    if (KIND == K_LAYER_NORM || KIND == K_RMS_NORM) { // CPU path yields Float32 for half inputs
        if (i == AXB_BF16 && o == AXB_F32)
            return launch_cfg<bf16_t, float, (KIND == K_LAYER_NORM || KIND == K_RMS_NORM) ? KIND : K_LAYER_NORM>(p, ctx, s);
        if (i == AXB_F16 && o == AXB_F32)
            return launch_cfg<f16_t, float, (KIND == K_LAYER_NORM || KIND == K_RMS_NORM) ? KIND : K_LAYER_NORM>(p, ctx, s);
    }
    return fail(AXB_ERR_UNSUPPORTED, "row kernel: dtype combination in=%d out=%d has no GPU kernel (fp32/bf16/fp16 only)", i, o);
}

axb_status run(int kind, const axb_tensor *in, const axb_tensor *w, const axb_tensor *b, int axis, float eps,
               const axb_tensor *out, axb_stream_t stream, const char *what) {
    AXB_TRY(validate_tensor(in, what));
    AXB_TRY(validate_tensor(out, what));
    AXB_REQUIRE(in->ndim >= 1, AXB_ERR_SHAPE, "%s: rank-0 input", what);
    AXB_REQUIRE(in->ndim == out->ndim, AXB_ERR_SHAPE, "%s: out rank differs", what);
    for (int d = 0; d < in->ndim; ++d)
        AXB_REQUIRE(in->shape[d] == out->shape[d], AXB_ERR_SHAPE, "%s: out shape differs in dim %d", what, d);
    if (axis < 0)
        axis += in->ndim; // softmax(-1) / layer_norm(-1) normalise (cpu_operations.cpp:3074-3078, operations.cpp:1797-1801)
    AXB_REQUIRE(axis >= 0 && axis < in->ndim, AXB_ERR_INVALID, "%s: axis out of range", what);
    if (numel(*in) == 0)
        return AXB_OK;
    AXB_REQUIRE(in->shape[axis] < (1ll << 31), AXB_ERR_UNSUPPORTED, "%s: row too long", what);
    RowParams p;
    memset(&p, 0, sizeof p);
    p.in = tensor_ptr(*in);
    p.out = tensor_ptr(*out);
    p.n = (int)in->shape[axis];
    p.in_es = in->strides[axis];
    p.out_es = out->strides[axis];
    p.in_dtype = in->dtype;
    p.out_dtype = out->dtype;
    p.eps = eps;
    p.rows = 1;
    for (int d = in->ndim - 1; d >= 0; --d) {
        if (d == axis || in->shape[d] == 1)
            continue;
        p.kshape[p.nk] = (uint32_t)in->shape[d];
        p.in_kstride[p.nk] = in->strides[d];
        p.out_kstride[p.nk] = out->strides[d];
        p.rows *= in->shape[d];
        p.nk++;
    }
    const int64_t isz = (int64_t)dtype_size(in->dtype), osz = (int64_t)dtype_size(out->dtype);
    bool vec = p.in_es == isz && p.out_es == osz && (reinterpret_cast<uintptr_t>(p.in) % (4 * isz)) == 0 &&
               (reinterpret_cast<uintptr_t>(p.out) % (4 * osz)) == 0;
    for (int d = 0; d < p.nk && vec; ++d)
        vec = p.in_kstride[d] % (4 * isz) == 0 && p.out_kstride[d] % (4 * osz) == 0;
    p.vec_ok = vec ? 1 : 0;
    auto bind_vec = [&](const axb_tensor *t, const char **ptr, int64_t *es, int32_t *dt, const char *nm) -> axb_status {
        if (!t)
            return AXB_OK;
        AXB_TRY(validate_tensor(t, nm));
        AXB_REQUIRE(t->dtype == AXB_F32 || t->dtype == AXB_BF16 || t->dtype == AXB_F16, AXB_ERR_UNSUPPORTED,
                    "%s: weight/bias dtype %d unsupported", what, t->dtype);
        // weight/bias broadcast against the input: only the axis extent may be > 1
        int64_t stride = 0;
        for (int d = 0; d < t->ndim; ++d) {
            int id = d + (in->ndim - t->ndim);
            if (t->shape[d] == 1)
                continue;
            AXB_REQUIRE(id == axis && t->shape[d] == in->shape[axis], AXB_ERR_SHAPE,
                        "%s: weight/bias must have shape {%lld} along the normalised axis", what, (long long)in->shape[axis]);
            stride = t->strides[d];
        }
        *ptr = tensor_ptr(*t);
        *es = stride;
        *dt = t->dtype;
        return AXB_OK;
    };
    AXB_TRY(bind_vec(w, &p.w, &p.w_es, &p.w_dtype, "weight"));
    AXB_TRY(bind_vec(b, &p.b, &p.b_es, &p.b_dtype, "bias"));
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    axb_status cst = AXB_OK;
    if (try_launch_cluster(kind, p, ctx, s, &cst)) { // wide rows with affine parameters: SM-pair cluster kernel
        last_path() = 2;
        return cst;
    }
    switch (kind) {
    case K_SOFTMAX: return dispatch<K_SOFTMAX>(p, ctx, s);
    case K_LOG_SOFTMAX: return dispatch<K_LOG_SOFTMAX>(p, ctx, s);
    case K_LAYER_NORM: return dispatch<K_LAYER_NORM>(p, ctx, s);
    default: return dispatch<K_RMS_NORM>(p, ctx, s);
    }
}

} // namespace
} // namespace rown
} // namespace axb

using namespace axb;

extern "C" {

axb_status axb_softmax(const axb_tensor *in, int axis, int is_log, const axb_tensor *out, axb_stream_t stream) {
    if (in && out)
        AXB_REQUIRE(in->dtype == out->dtype, AXB_ERR_INVALID, "axb_softmax: out dtype must equal the input dtype");
    return rown::run(is_log ? rown::K_LOG_SOFTMAX : rown::K_SOFTMAX, in, nullptr, nullptr, axis, 0.0f, out, stream,
                     "axb_softmax");
}

axb_status axb_layer_norm(const axb_tensor *in, const axb_tensor *weight, const axb_tensor *bias, int axis, float eps,
                          const axb_tensor *out, axb_stream_t stream) {
    return rown::run(rown::K_LAYER_NORM, in, weight, bias, axis, eps, out, stream, "axb_layer_norm");
}

axb_status axb_rms_norm(const axb_tensor *in, const axb_tensor *weight, int axis, float eps, const axb_tensor *out,
                        axb_stream_t stream) {
    return rown::run(rown::K_RMS_NORM, in, weight, nullptr, axis, eps, out, stream, "axb_rms_norm");
}

int axb_rownorm_last_path(void) { return axb::rown::last_path(); }

} // extern "C"
