// rownorm_cluster.cu — layer_norm / rms_norm WITH affine parameters on wide contiguous rows: thread-block-cluster kernel.
//
// Why: at n = 32 768 the weight and bias vectors (2 x 128 KiB as fp32) fit neither the registers nor the shared memory
// left beside a row buffer, so the single-CTA wide kernel (rownorm_wide.cuh) re-reads them from L2 for EVERY row:
// 8192 rows x 256 KiB = 2.1 GB of L2->SM traffic next to 1 GB (fp32) / 0.5 GB (bf16) of input. ncu r1: the L2 slices cap
// at ~11 TB/s, which held C3 layer_norm at 5.7 TB/s (fp32) and 3.0 TB/s (bf16 rows, fp32 parameters).
//
// Here a cluster of 2 CTAs (an SM pair, same GPC) splits the COLUMNS of every row: each CTA keeps its half of weight and
// bias resident in shared memory for the whole kernel (converted to fp32 once, laid out for conflict-free LDS.128), so
// the parameters cost one read per CTA instead of one per row. Row halves stream through per-warp 1-D bulk copies (TMA)
// exactly as in the wide kernel; the two row statistics (sum, then centred sum of squares) are reduced inside the CTA
// with shuffles and completed ACROSS the pair through distributed shared memory: each CTA stores its partial into both
// CTAs' exchange slots with st.async, which completes bytes on the destination's mbarrier (no cluster fences). Partials
// are added in rank order, so both CTAs get bit-identical statistics. Short half-rows process R = 2 rows per pass, which
// halves the number of exchanges per row.
//
// Numerics: identical op sequence to the other row kernels (see rownorm.cu header); reference: ops::layer_norm /
// rms_norm src/tensor/operations.cpp:1775-1858.
#include "rownorm_common.cuh"
#include "rownorm_wide.cuh"

namespace axb {
namespace rown {

constexpr int kClusterSize = 2;
// 512 threads x up to 128 registers: each thread keeps up to 64 row values (2 rows x 32), which halves the per-pass
// bookkeeping (reduction trees, barrier waits, address arithmetic — ~400 instructions per warp and pass, measured with
// ncu source counters) relative to 1024 threads x 32 values; the float work per element is unchanged.
constexpr int kClThreads = 512;
constexpr int kClWarps = kClThreads / 32;

__device__ __forceinline__ uint32_t cl_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cl_map(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Remote store that signals the destination CTA's mbarrier itself (complete_tx): data and notification travel together,
// so no cluster-scope release/acquire fence is needed. (A first version used st.shared::cluster + mbarrier.arrive
// .release.cluster / try_wait.acquire.cluster: ptxas turns those into MEMBAR.ALL.GPU — which waits for the thread's
// outstanding global stores of the previous row — and a CCTL.IVALL per waiting thread; measured 4.1 TB/s vs 5.7 TB/s.)
__device__ __forceinline__ void cl_store_signal_f32(uint32_t cluster_addr, float v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr),
                 "r"(__float_as_uint(v)), "r"(cluster_bar)
                 : "memory");
}
__device__ __forceinline__ void cl_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <typename W> __device__ __forceinline__ float wb_to_f32(const char *p);
template <> __device__ __forceinline__ float wb_to_f32<float>(const char *p) { return *reinterpret_cast<const float *>(p); }
template <> __device__ __forceinline__ float wb_to_f32<bf16_t>(const char *p) { return bf2f(*reinterpret_cast<const uint16_t *>(p)); }
template <> __device__ __forceinline__ float wb_to_f32<f16_t>(const char *p) { return h2f(*reinterpret_cast<const uint16_t *>(p)); }

// J = 16-byte vectors per lane per row-slice, R = rows per pass, EXACT: every lane owns J valid vectors (no masks). p.nv / p.vpw describe ONE slice (row / kClusterSize).
template <typename In, typename Out, typename W, int KIND, int J, int R, bool EXACT>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kClThreads, 1)
    rowcluster_kernel(const __grid_constant__ RowParams p) {
    constexpr int E = Vec16<In>::E;
    constexpr int V = J * E;
    constexpr int PCS = E / 4; // float4 pieces per vector of parameters
    extern __shared__ __align__(128) unsigned char cl_smem[];
    __shared__ __align__(8) uint64_t full_bar[kWideMaxStages * kClWarps];
    __shared__ __align__(8) uint64_t xbar[2];
    __shared__ float xchg[2][kClusterSize][R];
    __shared__ float scratch[2][R][kClWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cl_rank();
    const int64_t cluster_id = blockIdx.x / kClusterSize, n_clusters = gridDim.x / kClusterSize;
    const int S = p.stages;
    const int slice_cols = p.nv * E;
    const uint32_t slice_bytes = (uint32_t)p.nv * 16u;
    const int64_t slice_off_in = (int64_t)rank * slice_bytes;                           // bytes into a row
    const int64_t slice_off_out = (int64_t)rank * slice_cols * (int64_t)sizeof(Out);
    const int v0 = warp * p.vpw;
    const int v1 = min(v0 + p.vpw, p.nv);
    const uint32_t chunk_bytes = v1 > v0 ? (uint32_t)(v1 - v0) * 16u : 0u;
    float *wsm = reinterpret_cast<float *>(cl_smem);
    float *bsm = wsm + slice_cols;
    const uint32_t wsm_a = w_smem_u32(wsm), bsm_a = w_smem_u32(bsm);
    const uint32_t rows_base = w_smem_u32(cl_smem) + 2u * (uint32_t)slice_cols * 4u;
    const int64_t n_groups = (p.rows + R - 1) / R;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S * kClWarps; ++i)
            w_mbar_init(w_smem_u32(&full_bar[i]), 1);
        w_mbar_init(w_smem_u32(&xbar[0]), 1); // armed by this CTA's thread 0 with the expected byte count each use
        w_mbar_init(w_smem_u32(&xbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto offsets = [&](int64_t row, int64_t *in_off, int64_t *out_off) {
        if (p.nk <= 1) {
            *in_off = row * p.in_kstride[0];
            *out_off = row * p.out_kstride[0];
        } else {
            row_offsets(p, row, in_off, out_off);
        }
    };
    auto issue = [&](int64_t group, int stage) { // one lane per warp: R row-slices of this warp's chunk
        const uint32_t bar = w_smem_u32(&full_bar[stage * kClWarps + warp]);
        w_mbar_expect_tx(bar, chunk_bytes * R);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = min(group * R + r, p.rows - 1); // a ragged last group re-reads the last row
            int64_t in_off, out_off;
            offsets(row, &in_off, &out_off);
            w_bulk_load(rows_base + (uint32_t)(stage * R + r) * slice_bytes + (uint32_t)v0 * 16u,
                        p.in + in_off + slice_off_in + (int64_t)v0 * 16, chunk_bytes, bar);
        }
    };
    if (lane == 0 && chunk_bytes) {
        for (int s = 0; s < S; ++s) {
            const int64_t g = cluster_id + (int64_t)s * n_clusters;
            if (g < n_groups)
                issue(g, s);
        }
    }
    // resident parameters: element e of the slice -> vector vi = e / E, piece h, component k; stored so that lane L's
    // LDS.128 of piece h lands at ((vi/32)*PCS + h)*128 + (vi%32)*4 floats: consecutive lanes, consecutive 16 bytes.
    bool bad = false;
    for (int e = threadIdx.x; e < slice_cols; e += kClThreads) {
        const int vi = e / E, h = (e % E) / 4, k = e % 4;
        const int dst = (((vi >> 5) * PCS + h) * 32 + (vi & 31)) * 4 + k;
        const int64_t col = (int64_t)rank * slice_cols + e;
        const float wv = p.w ? wb_to_f32<W>(p.w + col * (int64_t)sizeof(W)) : 1.0f;
        const float bv = (KIND == K_LAYER_NORM && p.b) ? wb_to_f32<W>(p.b + col * (int64_t)sizeof(W)) : 0.0f;
        wsm[dst] = wv;
        bsm[dst] = bv;
        bad = bad || !(fabsf(wv) <= 3.0e38f) || !(fabsf(bv) <= 3.0e38f);
    }
    const bool wb_nonfinite = __syncthreads_or(bad) != 0;
    cl_sync_all(); // peers' mbarriers are initialised before anyone arrives on them remotely

    const float n_f = (float)p.n;
    const float rcp_n = __fdiv_rn(1.0f, n_f);
    // s / n, correctly rounded for all but pathological operands: quotient estimate + one exact-remainder correction
    auto div_n = [&](float s) {
        const float q = __fmul_rn(s, rcp_n);
        return __fmaf_rn(__fmaf_rn(-q, n_f, s), rcp_n, q);
    };
    // 1 / sqrt(v) to ~1 ulp: MUFU.RSQ + one Newton step (the reference divides by sqrt(v); the product form is already
    // covered by the 1e-5 contract)
    auto rsqrt1 = [&](float v) {
        const float r = rsqrtf(v);
        const float h = __fmul_rn(0.5f, r);
        return __fmaf_rn(h, __fmaf_rn(-__fmul_rn(v, r), r, 1.0f), r);
    };
    int flip = 0;
    uint32_t xk = 0; // exchange counter: slot = xk & 1, phase parity = (xk >> 1) & 1
    // CTA reduce (shuffles + one smem hop) then pair reduce through DSMEM; every thread of both CTAs gets the same sums
    auto allreduce = [&](float (&v)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1)
                v[r] += __shfl_xor_sync(0xFFFFFFFFu, v[r], d);
            if (lane == 0)
                scratch[flip][r][warp] = v[r];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float t = scratch[flip][r][lane & (kClWarps - 1)];
#pragma unroll
            for (int d = kClWarps / 2; d >= 1; d >>= 1)
                t += __shfl_xor_sync(0xFFFFFFFFu, t, d);
            v[r] = t;
        }
        flip ^= 1;
        const uint32_t slot = xk & 1u, parity = (xk >> 1) & 1u;
        ++xk;
        if (threadIdx.x == 0)
            w_mbar_expect_tx(w_smem_u32(&xbar[slot]), kClusterSize * R * 4u);
        if (threadIdx.x < kClusterSize) { // thread t delivers this CTA's partials to CTA t of the pair (itself included)
            const uint32_t dst = cl_map(w_smem_u32(&xchg[slot][rank][0]), threadIdx.x);
            const uint32_t bar = cl_map(w_smem_u32(&xbar[slot]), threadIdx.x);
#pragma unroll
            for (int r = 0; r < R; ++r)
                cl_store_signal_f32(dst + 4u * r, v[r], bar);
        }
        w_mbar_wait(w_smem_u32(&xbar[slot]), parity);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float t = xchg[slot][0][r];
#pragma unroll
            for (int c = 1; c < kClusterSize; ++c)
                t += xchg[slot][c][r];
            v[r] = t;
        }
    };

    int stage = 0;
    uint32_t parity = 0;
    for (int64_t g = cluster_id; g < n_groups; g += n_clusters) {
        float x[R][V];
        if (chunk_bytes)
            w_mbar_wait(w_smem_u32(&full_bar[stage * kClWarps + warp]), parity);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t buf = rows_base + (uint32_t)(stage * R + r) * slice_bytes;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int vi = v0 + j * 32 + lane;
                if (EXACT || vi < v1) {
                    Vec16<In>::unpack(w_lds128(buf + (uint32_t)vi * 16u), &x[r][j * E]);
                } else {
#pragma unroll
                    for (int k = 0; k < E; ++k)
                        x[r][j * E + k] = 0.0f;
                }
            }
        }
        // The stage is refilled only after its reads have HAPPENED (w_loads_landed, rownorm_wide.cuh: issuing the LDS
        // is not enough, a refill from an L2-resident input can overtake them), and after the first sweep over the
        // row's values so that the wait for the data is not exposed.
        auto refill = [&]() {
            uint32_t landed = 0;
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < J; ++j)
                    landed |= __float_as_uint(x[r][j * E]);
            w_loads_landed(landed);
            __syncwarp();
            if (lane == 0 && chunk_bytes) {
                const int64_t ng = g + (int64_t)S * n_clusters;
                if (ng < n_groups)
                    issue(ng, stage);
            }
            if (++stage == S) {
                stage = 0;
                parity ^= 1u;
            }
        };

        float scale[R];
        bool exact_row[R]; // NaN/inf statistics: every output of the row is NaN -> use the reference's NaN encoding
        if (KIND == K_LAYER_NORM) {
            float s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                for (int i = 0; i < V; i += 4) {
                    s0 += x[r][i]; s1 += x[r][i + 1]; s2 += x[r][i + 2]; s3 += x[r][i + 3];
                }
                s[r] = (s0 + s1) + (s2 + s3);
            }
            refill();
            allreduce(s);
            float q[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float mean = round_to<In>(div_n(s[r]));
                float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const bool ok = EXACT || (v0 + j * 32 + lane) < v1;
#pragma unroll
                    for (int k = 0; k < E; k += 4) { // centred value kept (exact subtraction); square-accumulate fused
                        float *c = &x[r][j * E + k];
                        c[0] = ok ? __fsub_rn(c[0], mean) : 0.0f; q0 = __fmaf_rn(c[0], c[0], q0);
                        c[1] = ok ? __fsub_rn(c[1], mean) : 0.0f; q1 = __fmaf_rn(c[1], c[1], q1);
                        c[2] = ok ? __fsub_rn(c[2], mean) : 0.0f; q2 = __fmaf_rn(c[2], c[2], q2);
                        c[3] = ok ? __fsub_rn(c[3], mean) : 0.0f; q3 = __fmaf_rn(c[3], c[3], q3);
                    }
                }
                q[r] = (q0 + q1) + (q2 + q3);
            }
            allreduce(q);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                scale[r] = rsqrt1(__fadd_rn(div_n(q[r]), p.eps));
                exact_row[r] = !(fabsf(scale[r]) <= 3.0e38f); // NaN mean/variance end up here as well
            }
        } else {
            float q[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float q0 = 0.0f, q1 = 0.0f, q2 = 0.0f, q3 = 0.0f;
#pragma unroll
                for (int i = 0; i < V; i += 4) {
                    if constexpr (sizeof(In) == 4) {
                        q0 = __fmaf_rn(x[r][i], x[r][i], q0);
                        q1 = __fmaf_rn(x[r][i + 1], x[r][i + 1], q1);
                        q2 = __fmaf_rn(x[r][i + 2], x[r][i + 2], q2);
                        q3 = __fmaf_rn(x[r][i + 3], x[r][i + 3], q3);
                    } else { // the CPU path squares in the half type (bfloat16.hpp:55-66): round each product
                        q0 += round_to<In>(__fmul_rn(x[r][i], x[r][i]));
                        q1 += round_to<In>(__fmul_rn(x[r][i + 1], x[r][i + 1]));
                        q2 += round_to<In>(__fmul_rn(x[r][i + 2], x[r][i + 2]));
                        q3 += round_to<In>(__fmul_rn(x[r][i + 3], x[r][i + 3]));
                    }
                }
                q[r] = (q0 + q1) + (q2 + q3);
            }
            refill();
            allreduce(q);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float ms = round_to<In>(div_n(q[r]));
                scale[r] = rsqrt1(__fadd_rn(ms, p.eps));
                exact_row[r] = !(fabsf(scale[r]) <= 3.0e38f);
            }
        }

        // y = (xc * (1/sd)) * w + b: parameters come from shared memory, once per vector for all R rows; absent weight /
        // bias were materialised as 1 / 0 so the inner loop is branch-free (one FMUL + one FFMA per element).
        // bf16 results use the hardware converter (NaN -> 0x7FFF) unless the row's statistics or this CTA's parameters
        // are non-finite — the only ways a NaN can appear (|xc/sd| <= sqrt(n), so nothing overflows) — in which case
        // the reference's own encoding (NaN -> 0x7FC0, bfloat16.hpp:16-27) is produced by the software conversion.
        constexpr bool HWPACK = std::is_same<Out, bf16_t>::value;
        char *dst[R];
        bool row_ok[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = g * R + r;
            row_ok[r] = row < p.rows;
            int64_t in_off, out_off;
            offsets(row_ok[r] ? row : 0, &in_off, &out_off);
            dst[r] = p.out + out_off + slice_off_out;
            exact_row[r] = exact_row[r] || wb_nonfinite;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int vi = v0 + j * 32 + lane;
            if (EXACT || vi < v1) {
                float wv[E], bv[E];
                const uint32_t po = (uint32_t)((((vi >> 5) * PCS) * 32 + (vi & 31)) * 16);
#pragma unroll
                for (int h = 0; h < PCS; ++h) {
                    const uint4 qw = w_lds128(wsm_a + po + (uint32_t)h * 512u);
                    wv[4 * h] = __uint_as_float(qw.x); wv[4 * h + 1] = __uint_as_float(qw.y);
                    wv[4 * h + 2] = __uint_as_float(qw.z); wv[4 * h + 3] = __uint_as_float(qw.w);
                    if (KIND == K_LAYER_NORM) {
                        const uint4 qb = w_lds128(bsm_a + po + (uint32_t)h * 512u);
                        bv[4 * h] = __uint_as_float(qb.x); bv[4 * h + 1] = __uint_as_float(qb.y);
                        bv[4 * h + 2] = __uint_as_float(qb.z); bv[4 * h + 3] = __uint_as_float(qb.w);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float y[E];
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        const float t = __fmul_rn(x[r][j * E + k], scale[r]);
                        y[k] = KIND == K_LAYER_NORM ? __fmaf_rn(t, wv[k], bv[k]) : __fmul_rn(t, wv[k]);
                    }
                    if (row_ok[r]) {
                        char *d = dst[r] + (int64_t)vi * (E * sizeof(Out));
                        if (HWPACK && !exact_row[r])
                            store_vec<Out, E, true>(d, y);
                        else
                            store_vec<Out, E, false>(d, y);
                    }
                }
            }
        }
    }
    cl_sync_all(); // no CTA leaves while its peer may still deliver into its shared memory
}

namespace {

template <typename In, typename Out, typename W, int KIND>
bool launch_cluster_typed(RowParams p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    constexpr int E = 16 / (int)sizeof(In);
    const int64_t isz = sizeof(In), osz = sizeof(Out);
    if (!p.w && !p.b)
        return false; // nothing to keep resident: the single-CTA wide kernel is the right shape
    if (p.in_es != isz || p.out_es != osz || ((int64_t)p.n * isz) % (16 * kClusterSize) != 0)
        return false;
    const int64_t nv = (int64_t)p.n * isz / 16 / kClusterSize; // vectors per row slice
    if (nv < 512 || nv % 32 != 0 || (int64_t)p.n * 8 < 64 * 1024 || p.rows < 4 * (int64_t)ctx->sm_count)
        return false; // narrow rows: parameters stay in L1/L2; few rows: loading the resident copy would not amortise
    auto al16 = [](const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!al16(p.in) || !al16(p.out))
        return false;
    for (int d = 0; d < p.nk; ++d)
        if (p.in_kstride[d] % 16 != 0 || p.out_kstride[d] % 16 != 0)
            return false;
    auto wb_ok = [&](const char *q, int64_t es, int dt) { return !q || es == (int64_t)dtype_size(dt); };
    if (!wb_ok(p.w, p.w_es, p.w_dtype) || !wb_ok(p.b, p.b_es, p.b_dtype))
        return false;
    static const bool disabled = getenv("AXB_NO_CLUSTER_ROWS") != nullptr; // A/B switch for profiling
    if (disabled)
        return false;
    p.nv = (int)nv;
    p.vpw = (int)((nv + kClWarps - 1) / kClWarps);
    const int j_need = (p.vpw + 31) / 32;
    constexpr int JMAX = 32 / E;
    if (j_need > JMAX)
        return false;
    const size_t slice_bytes = (size_t)nv * 16;
    const size_t wb_bytes = 2 * (size_t)nv * E * 4;
    const size_t budget = ctx->smem_optin - 4096; // static: mbarriers, exchange slots, reduction scratch
    if (wb_bytes + slice_bytes > budget)
        return false;
    // rows per pass: two when the registers (<= 64 resident values per thread) and shared memory allow it
    int jt = j_need <= 1 ? 1 : j_need <= 2 ? 2 : j_need <= 4 ? 4 : 8;
    int R = (2 * jt * E <= 64 && wb_bytes + 2 * slice_bytes <= budget) ? 2 : 1;
    static const char *force_r = getenv("AXB_CLUSTER_ROWS_PER_PASS");
    if (force_r && atoi(force_r) == 1)
        R = 1;
    const size_t per_stage = slice_bytes * R;
    size_t stages = std::max<size_t>(1, (64 * 1024 + per_stage - 1) / per_stage); // >= 64 KiB in flight per SM
    stages = std::min<size_t>({stages, (size_t)kWideMaxStages, (budget - wb_bytes) / per_stage});
    p.stages = (int)stages;
    const size_t dyn = wb_bytes + stages * per_stage;
    const int64_t n_groups = (p.rows + R - 1) / R;
    const unsigned grid = (unsigned)(kClusterSize * std::min<int64_t>(n_groups, ctx->sm_count / kClusterSize));
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "cluster row kernel: cannot reserve %zu bytes of shared memory: %s", dyn, cudaGetErrorString(e));
            return;
        }
        kern<<<grid, kClThreads, dyn, s>>>(p);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) {
            *st = fail(AXB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(le), __FILE__, __LINE__);
            return;
        }
        count_launch();
        *st = AXB_OK;
    };
    // instantiated shapes: J in {2, 4, 8} (fp32) / {1, 2, 4} (half) vectors per lane, R = 2 where the budget allows;
    // the mask-free variant exists for the widest shape (n = 32 768)
    constexpr int JA = JMAX, JB = JMAX / 2, JC = JMAX / 4;
    const int jsel = jt > JB ? JA : jt > JC ? JB : JC;
    const bool exact = p.vpw == JA * 32 && nv == (int64_t)kClWarps * p.vpw;
    if (R == 2) {
        if (exact)
            go(rowcluster_kernel<In, Out, W, KIND, JA, 2, true>);
        else if (jsel == JA)
            go(rowcluster_kernel<In, Out, W, KIND, JA, 2, false>);
        else if (jsel == JB)
            go(rowcluster_kernel<In, Out, W, KIND, JB, 2, false>);
        else
            go(rowcluster_kernel<In, Out, W, KIND, JC, 2, false>);
    } else {
        if (exact)
            go(rowcluster_kernel<In, Out, W, KIND, JA, 1, true>);
        else if (jsel == JA)
            go(rowcluster_kernel<In, Out, W, KIND, JA, 1, false>);
        else if (jsel == JB)
            go(rowcluster_kernel<In, Out, W, KIND, JB, 1, false>);
        else
            go(rowcluster_kernel<In, Out, W, KIND, JC, 1, false>);
    }
    return true;
}

template <typename In, typename Out, int KIND> bool pick_w(const RowParams &p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    const int wdt = p.w ? p.w_dtype : p.b_dtype;
    if (p.w && p.b && p.w_dtype != p.b_dtype)
        return false;
    if (wdt == AXB_F32)
        return launch_cluster_typed<In, Out, float, KIND>(p, ctx, s, st);
    if (wdt == p.in_dtype)
        return launch_cluster_typed<In, Out, In, KIND>(p, ctx, s, st);
    return false;
}

template <int KIND> bool pick_io(const RowParams &p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    const int i = p.in_dtype, o = p.out_dtype;
    if (i == AXB_F32 && o == AXB_F32)
        return pick_w<float, float, KIND>(p, ctx, s, st);
    if (i == AXB_BF16 && o == AXB_BF16)
        return pick_w<bf16_t, bf16_t, KIND>(p, ctx, s, st);
    if (i == AXB_BF16 && o == AXB_F32)
        return pick_w<bf16_t, float, KIND>(p, ctx, s, st);
    if (i == AXB_F16 && o == AXB_F16)
        return pick_w<f16_t, f16_t, KIND>(p, ctx, s, st);
    if (i == AXB_F16 && o == AXB_F32)
        return pick_w<f16_t, float, KIND>(p, ctx, s, st);
    return false;
}

} // namespace

bool try_launch_cluster(int kind, const RowParams &p, DeviceCtx *ctx, cudaStream_t s, axb_status *st) {
    if (kind == K_LAYER_NORM)
        return pick_io<K_LAYER_NORM>(p, ctx, s, st);
    if (kind == K_RMS_NORM)
        return pick_io<K_RMS_NORM>(p, ctx, s, st);
    return false;
}

} // namespace rown
} // namespace axb
