// gemm_tcgen05.cu — kernel family 3: bf16 / fp16 matmul and batched matmul on the 5th-generation tensor cores.
//
// Replaces the BLAS call of the reference's matmul path (CPUMatMulOperation -> matmul_2d -> BlasBackend::sgemm,
// src/backends/cpu/cpu_operations.cpp:1773-1883, :2028-2053 for the batch loop) for half-precision operands; the
// reference has no BLAS path for bf16 at all (tiled triple loop with a bf16 accumulator, :1835-1882), so the numerics
// contract is the north_star's: fp32 accumulate, ONE rounding to the output type, <= 1e-2 relative vs fp32.
//
// Design (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer : cp.async.bulk.tensor (3-D maps: inner, rows, batch) into a 4-stage shared-memory ring,
//                              128-byte swizzle, completion on mbarriers (expect_tx)
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16),
//                              A/B straight from shared memory through UMMA descriptors, fp32 accumulators in TMEM;
//                              tcgen05.commit releases ring slots / publishes finished accumulators
//   warps 2..9  epilogue     : tcgen05.ld (32 lanes x 32 columns per warp) -> bias / activation -> one rounding ->
//                              128-bit global stores (two warps per TMEM lane quarter, alternate column chunks). TMEM holds TWO 128x256 fp32 accumulators (all 512 columns), so the
//                              epilogue of tile i overlaps the MMAs of tile i+1.
// Transposes are never materialised: an operand whose contraction dim is contiguous is fed K-major, one whose M/N dim
// is contiguous is fed MN-major (instruction-descriptor major bits + the matching UMMA canonical layout), which covers
// x.T() views and nn::Linear's transpose_b. Ragged M/N/K use TMA's out-of-bounds zero fill and masked stores.
// Anything TMA cannot describe (unaligned base/strides, inner stride != 1, non-collapsible batch) returns to the SIMT path.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "driver.h"
#include "gemm.h"
#include "hwy_math.cuh"
#include "reduce_common.cuh"

namespace axb {
namespace gemm {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
constexpr int B_BYTES = BN * BK * 2;        // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_BYTES = 8 * 4096;   // one 32 x 32-word tile per epilogue warp (row <-> column exchange for the stores)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + STAGING_BYTES;
constexpr int NUM_THREADS = 320; // TMA producer, MMA issuer, 8 epilogue warps
constexpr uint32_t TMEM_COLS = 512;

constexpr int kMaxTaps = 64;
struct TcParams {
    int64_t M, N, K, batch;
    int32_t tiles_m, tiles_n;
    int64_t num_tiles;
    char *c;
    int64_t c_rs, c_cs, c_bs;       // bytes
    const char *bias;
    int64_t bias_rs, bias_cs, bias_bs;
    int32_t bias_dtype, act, out_dtype;
    int32_t a_bmult, b_bmult;       // 0 when the operand is broadcast over the batch
    uint32_t idesc;
    // fp32-accurate mode (see try_split_f32): operands are [rows, 3*Kp] bf16 (hi | mid | lo pieces), the K loop walks
    // six (A piece, B piece) combinations of kbp k-blocks each
    int32_t split, kbp;
    int64_t Kp;
    int32_t bn; // tile width in N (64 / 128 / 256)
    int32_t group_m; // CTA-pair kernel: tile rows per band (see pair_tile_coords)
    // implicit-GEMM convolution (see conv_tc): conv_taps > 0 turns the K loop into (filter tap, 64-channel block) steps.
    // A is [tap][C_out][pieces * Cp] (3-D map, tap in the batch coordinate), B the phase-split channels-last input
    // [image * phases + phase][H'][W'][pieces * Cp] (4-D map); tap t reads B at column n0 + conv_qw[t], row oh + conv_qh[t],
    // plane image * conv_phases + conv_ph[t] — out-of-range coordinates are zero-filled by TMA, which IS the padding.
    // "batch" counts (image, output row group) pairs: b = image * conv_oh + group; a group is conv_rows consecutive output
    // rows, each filling one 64-column chunk of the tile (images narrower than 64 columns: a 256-wide tile holds four rows,
    // so the weights are fetched once per four rows instead of once per row); c_hs is the byte stride of an output row,
    // conv_ohs the number of output rows of an image.
    int32_t conv_taps, conv_cblocks, conv_oh, conv_phases, conv_rows, conv_ohs;
    // conv_rows > 1: output rows per 64-column chunk of the tile (1: a chunk is one row of up to 64 columns; 2: images at most 32
    // columns wide, a chunk is a {2 rows x 32 columns} TMA box, i.e. twice the rows per tile for the same MMA work)
    int32_t conv_rpc;
    int64_t c_hs;
    int16_t conv_qw[kMaxTaps], conv_qh[kMaxTaps], conv_ph[kMaxTaps];
};

// piece (0 = hi, 1 = mid, 2 = lo) of A and B for combination c, smallest products first
__device__ __forceinline__ int split_piece_a(int c) { return (0x001120 >> (4 * c)) & 0xF; } // c = 0..5: 0,2,1,1,0,0
__device__ __forceinline__ int split_piece_b(int c) { return (0x010102 >> (4 * c)) & 0xF; } // c = 0..5: 2,0,1,0,1,0

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One lane of a fully converged warp. The MMA-issuing warp runs its schedule warp-uniformly and only the tcgen05
// instructions sit under this predicate: descriptors and barrier addresses then live in uniform registers. Issuing from
// inside `if (lane == 0)` cost ~10 extra instructions per MMA (R2UR moves + an ELECT retry loop around every UTCHMMA,
// profiles/r2_flash_attention_notes.txt) — as long as a 128x128x16 MMA runs.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

// GELU of an fp32 GEMM result. The eight epilogue warps of a CTA have to push a whole 128 x BN tile through the
// activation while the tensor cores work on the next one; with the bit-exact Highway sequence (hwym::Gelu, ~60
// instructions per element) the (12000 x 384) x (384 x 1536) MLP GEMM of the whisper encoder spent more time in the
// epilogue than in the MMAs (114 us against 67 us for the same FLOPs without an activation). The matmul result feeding
// it is tolerance-level anyway (1e-5 contract), so this variant keeps the reference's formula — Abramowitz-Stegun
// 7.1.26, hwy_math_ext-inl.h:1-37 — and evaluates its reciprocal and exponential on the MUFU unit: ~16 instructions,
// |difference from hwym::Gelu| <= 3e-7 * max(1, |x|). Half-precision outputs keep hwym::Gelu (identical to the unfused op).
__device__ __forceinline__ float gelu_epilogue_f32(float x) {
    const float p = 0.3275911f, a1 = 0.254829592f, a2 = -0.284496736f, a3 = 1.421413741f, a4 = -1.453152027f,
                a5 = 1.061405429f;
    const float s = x * 0.7071067811865476f;
    const float as = fabsf(s);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(p, as, 1.0f)));
    float poly = fmaf(t, a5, a4);
    poly = fmaf(t, poly, a3);
    poly = fmaf(t, poly, a2);
    poly = fmaf(t, poly, a1);
    poly *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(s * s * -1.4426950408889634f));
    const float r = fmaf(-poly, e, 1.0f);     // erf(|s|); NaN stays NaN, |s| = inf gives 1
    const float half_x = 0.5f * x;
    return fmaf(half_x, copysignf(r, s), half_x); // 0.5 * x * (1 + erf(s))
}
__device__ __forceinline__ float ld_bias(const char *p, int dt) {
    return dt == AXB_F32 ? *reinterpret_cast<const float *>(p)
                         : dt == AXB_BF16 ? red::bf2f(*reinterpret_cast<const uint16_t *>(p))
                                          : red::h2f(*reinterpret_cast<const uint16_t *>(p));
}

// One 128 x BN_ accumulator tile: TMEM -> bias / activation -> one rounding -> global memory. Shared by the one-CTA and
// the CTA-pair kernel (q = TMEM lane quarter of the warp, half = which alternate 32-column chunks it takes, ewarp = index
// of the warp's staging tile).
template <int BN_>
__device__ __forceinline__ void epilogue_tile(const TcParams &p, int64_t b, int m0, int n0, uint32_t taddr, bool grouped, int q,
                                              int half, int lane, int ewarp, uint32_t staging_base) {
    const bool out_f32 = p.out_dtype == AXB_F32, out_bf16 = p.out_dtype == AXB_BF16;
    const int osz = out_f32 ? 4 : 2;
    // convolution: b = image * conv_oh + output row (see TcParams)
    const int64_t conv_oh0 = p.conv_taps ? (b % p.conv_oh) * p.conv_rows : 0; // first output row of this tile's group
    const int64_t cbase0 = p.conv_taps ? (b / p.conv_oh) * p.c_bs + conv_oh0 * p.c_hs : b * p.c_bs;
    const bool c_vec = (p.c_cs == osz) && ((reinterpret_cast<uintptr_t>(p.c) | (uintptr_t)p.c_rs | (uintptr_t)p.c_bs | (uintptr_t)p.c_hs) % 16 == 0);
    // bias as a row vector shared by all rows (nn::Linear): read once per chunk with vector loads
    const int bsz = p.bias_dtype == AXB_F32 ? 4 : 2;
    const bool bias_vec = p.bias && p.bias_rs == 0 && p.bias_cs == bsz && ((reinterpret_cast<uintptr_t>(p.bias) | (uintptr_t)p.bias_bs) % 16 == 0);
    auto round_out = [&](float (&f)[32]) { // round to the half output type and come back (hardware converters)
        if (out_bf16) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                uint32_t h;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f[i + 1]), "f"(f[i]));
                f[i] = __uint_as_float(h << 16);
                f[i + 1] = __uint_as_float(h & 0xFFFF0000u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                f[i] = red::h2f(red::f2h(f[i]));
        }
    };
    const int64_t m = m0 + q * 32 + lane;
    const char *brow = p.bias ? p.bias + b * p.bias_bs + m * p.bias_rs : nullptr;
#pragma unroll 1
    for (int c = half; c < BN_ / 32; c += 2) {
        int64_t nbase = n0 + c * 32, cbase = cbase0;
        // rpc4: this 32-column segment is TWO output rows of at most 16 columns (columns 0..15 -> row r, 16..31 -> row r + 1)
        const bool rpc4 = p.conv_rows > 1 && p.conv_rpc == 4;
        int64_t rows_left = 1;
        if (p.conv_rows > 1) { // 64-column chunk k of the tile holds output rows conv_oh0 + k * rpc .. (rpc = 1: one 64-column row,
            // two 32-column segments; rpc = 2: two rows of 32 columns, one segment each; rpc = 4: four rows of 16, two per segment)
            const int r = rpc4 ? 2 * c : p.conv_rpc == 2 ? c : c >> 1;
            nbase = p.conv_rpc >= 2 ? 0 : (c & 1) * 32;
            rows_left = p.conv_ohs - (conv_oh0 + r);
            if (rows_left <= 0 || nbase >= p.N)
                continue; // warp-uniform
            cbase += r * p.c_hs;
        } else if (nbase >= p.N) {
            break; // warp-uniform
        }
        // byte offset (from the segment's first row) of column `col` of the segment, and whether it exists
        auto col_off = [&](int col, bool *ok) -> int64_t {
            if (rpc4) {
                const int sub = col >> 4, cl = col & 15;
                *ok = cl < p.N && sub < rows_left;
                return sub * p.c_hs + (int64_t)cl * p.c_cs;
            }
            *ok = nbase + col < p.N;
            return (nbase + col) * p.c_cs;
        };
        char *crow = p.c + cbase + m * p.c_rs;
        uint32_t v[32];
        tc_ld32(taddr + c * 32, v);
        tc_wait_ld();
        if (grouped) { // main (hi.hi) + small-products accumulator
            uint32_t w[32];
            tc_ld32(taddr + 128u + c * 32, w);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
        }
        const bool full = nbase + 32 <= p.N;
        // a ragged last chunk still takes the 128-bit exchange store when N is a whole number of 16-byte segments (the
        // segments past N are simply not written): N = 56 bf16 columns would otherwise leave 24 of them to 2-byte stores
        const bool vec_store = c_vec && (full || (p.N * osz) % 16 == 0);
        // columns a few elements apart and rows far apart (one phase of a transposed convolution writing out[.., rw :: sw]):
        // the same shared-memory exchange, then element stores with the lanes of a row on neighbouring columns
        const bool strided_store = !c_vec && p.c_cs > osz && p.c_cs <= 64 && p.c_rs > 32 * p.c_cs;
        if (m >= p.M && !vec_store && !strided_store)
            continue; // (the exchange-based stores below need the whole warp)
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i)
            f[i] = __uint_as_float(v[i]);
        // ---- bias: a separate op in the reference, applied to the matmul result rounded to its dtype ----------
        if (p.bias) {
            if (!out_f32)
                round_out(f);
            if (bias_vec && full) {
                const char *bp = p.bias + b * p.bias_bs + nbase * bsz;
                if (bsz == 4) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 w = __ldg(reinterpret_cast<const float4 *>(bp) + (i >> 2));
                        f[i] += w.x; f[i + 1] += w.y; f[i + 2] += w.z; f[i + 3] += w.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(bp) + (i >> 3));
                        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (p.bias_dtype == AXB_BF16) {
                                f[i + 2 * k] += __uint_as_float(ww[k] << 16);
                                f[i + 2 * k + 1] += __uint_as_float(ww[k] & 0xFFFF0000u);
                            } else {
                                f[i + 2 * k] += red::h2f((uint16_t)(ww[k] & 0xFFFFu));
                                f[i + 2 * k + 1] += red::h2f((uint16_t)(ww[k] >> 16));
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (rpc4 || nbase + i < p.N) // (rpc4: a convolution's bias is per output channel, bias_cs == 0)
                        f[i] += ld_bias(brow + (nbase + i) * p.bias_cs, p.bias_dtype);
            }
        }
        // ---- activation (also a separate op: its input is rounded to the output dtype first) -----------------
        if (p.act) {
            if (!out_f32)
                round_out(f);
            switch (p.act) {
            case AXB_OP_RELU:
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = f[i] > 0.0f ? f[i] : 0.0f;
                break;
            case AXB_OP_GELU:
                if (out_f32) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = gelu_epilogue_f32(f[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) f[i] = hwym::Gelu(f[i]);
                }
                break;
            case AXB_OP_SILU:
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = hwym::Silu(f[i]);
                break;
            case AXB_OP_SIGMOID:
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = hwym::Sigmoid(f[i]);
                break;
            default:
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = hwym::Tanh(f[i]);
                break;
            }
        }
        // ---- store -----------------------------------------------------------------------------------------
        if (out_f32) {
            if (vec_store) {
                // Each lane owns one ROW of the 32 x 32 chunk; written straight out, a warp store would touch 32
                // different 128-byte lines with 16 bytes each. The chunk goes through a swizzled shared-memory
                // tile instead so that 8 lanes cover one row: every store instruction writes 4 full lines.
                const uint32_t tile = staging_base + (uint32_t)ewarp * 4096u;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + lane * 128u + ((i ^ (lane & 7)) << 4)),
                                 "r"(__float_as_uint(f[4 * i])), "r"(__float_as_uint(f[4 * i + 1])), "r"(__float_as_uint(f[4 * i + 2])),
                                 "r"(__float_as_uint(f[4 * i + 3]))
                                 : "memory");
                __syncwarp();
                const int cw = lane & 7;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3);
                    uint4 w;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                                 : "r"(tile + rr * 128u + ((cw ^ (rr & 7)) << 4)));
                    const int64_t mm = (int64_t)m0 + q * 32 + rr;
                    bool ok;
                    const int64_t off = col_off(cw * 4, &ok);
                    if (mm < p.M && ok)
                        *reinterpret_cast<uint4 *>(p.c + cbase + mm * p.c_rs + off) = w;
                }
                __syncwarp();
            } else if (strided_store) {
                // written straight out, a warp store would touch 32 rows with 4 bytes each; after the exchange 8 lanes
                // cover 8 neighbouring columns of one row (2 sectors at a column stride of 8 bytes) and an instruction 4 rows
                const uint32_t tile = staging_base + (uint32_t)ewarp * 4096u;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + lane * 128u + ((i ^ (lane & 7)) << 4)),
                                 "r"(__float_as_uint(f[4 * i])), "r"(__float_as_uint(f[4 * i + 1])), "r"(__float_as_uint(f[4 * i + 2])),
                                 "r"(__float_as_uint(f[4 * i + 3]))
                                 : "memory");
                __syncwarp();
                const int cw = lane & 7;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3);
                    const int64_t mm = (int64_t)m0 + q * 32 + rr;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int col = k * 8 + cw;
                        uint32_t w;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w)
                                     : "r"(tile + rr * 128u + ((((uint32_t)col >> 2) ^ (rr & 7)) << 4) + ((uint32_t)col & 3u) * 4u));
                        bool ok;
                        const int64_t off = col_off(col, &ok);
                        if (mm < p.M && ok)
                            *reinterpret_cast<uint32_t *>(p.c + cbase + mm * p.c_rs + off) = w;
                    }
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    bool ok;
                    const int64_t off = col_off(i, &ok);
                    if (ok)
                        *reinterpret_cast<float *>(crow + off) = f[i];
                }
            }
            continue;
        }
        uint32_t h[16];
        if (out_bf16) {
            // hardware pack (NaN -> 0x7FFF); the reference encodes NaN as 0x7FC0 (bfloat16.hpp:16-27): chunks that
            // contain a NaN are re-packed in software (rare)
            float nan_acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(f[2 * i + 1]), "f"(f[2 * i]));
                asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(nan_acc) : "f"(f[2 * i]), "f"(f[2 * i + 1]));
            }
            if (nan_acc != nan_acc) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    h[i] = (uint32_t)red::f2bf(f[2 * i]) | ((uint32_t)red::f2bf(f[2 * i + 1]) << 16);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                h[i] = (uint32_t)red::f2h(f[2 * i]) | ((uint32_t)red::f2h(f[2 * i + 1]) << 16);
        }
        if (vec_store) {
            // same exchange for 64-byte rows: 4 lanes per row, 8 rows (4 full lines) per store instruction
            const uint32_t tile = staging_base + (uint32_t)ewarp * 4096u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + lane * 64u + ((i ^ ((lane >> 1) & 3)) << 4)),
                             "r"(h[4 * i]), "r"(h[4 * i + 1]), "r"(h[4 * i + 2]), "r"(h[4 * i + 3])
                             : "memory");
            __syncwarp();
            const int cw = lane & 3;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + (lane >> 2);
                uint4 w;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                             : "r"(tile + rr * 64u + ((cw ^ ((rr >> 1) & 3)) << 4)));
                const int64_t mm = (int64_t)m0 + q * 32 + rr;
                bool ok;
                const int64_t off = col_off(cw * 8, &ok);
                if (mm < p.M && ok)
                    *reinterpret_cast<uint4 *>(p.c + cbase + mm * p.c_rs + off) = w;
            }
            __syncwarp();
        } else if (strided_store) { // 64-byte rows: 4 lanes per row on neighbouring columns, 8 rows per instruction
            const uint32_t tile = staging_base + (uint32_t)ewarp * 4096u;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + lane * 64u + ((i ^ ((lane >> 1) & 3)) << 4)),
                             "r"(h[4 * i]), "r"(h[4 * i + 1]), "r"(h[4 * i + 2]), "r"(h[4 * i + 3])
                             : "memory");
            __syncwarp();
            const int cw = lane & 3;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + (lane >> 2);
                const int64_t mm = (int64_t)m0 + q * 32 + rr;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int col = k * 4 + cw;
                    uint16_t w;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(w)
                                 : "r"(tile + rr * 64u + ((((uint32_t)col >> 3) ^ ((rr >> 1) & 3)) << 4) + ((uint32_t)col & 7u) * 2u));
                    bool ok;
                    const int64_t off = col_off(col, &ok);
                    if (mm < p.M && ok)
                        *reinterpret_cast<uint16_t *>(p.c + cbase + mm * p.c_rs + off) = w;
                }
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                bool ok;
                const int64_t off = col_off(i, &ok);
                if (ok)
                    *reinterpret_cast<uint16_t *>(crow + off) = (uint16_t)((i & 1) ? (h[i >> 1] >> 16) : (h[i >> 1] & 0xFFFFu));
            }
        }
    }
}

// A_MN / B_MN: operand is MN-major (its M / N dim is the contiguous one) instead of K-major.
// BN_: tile width in N (64 / 128 / 256). The shared-memory ring keeps the 256-wide stage stride; narrower tiles simply
// load and multiply less of it (N = 64 attention values, N = 384 Linear layers: 3 exact 128-wide tiles instead of 2 x 256).
template <bool A_MN, bool B_MN, int BN_>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem = (raw + 1023u) & ~1023u; // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t bar_base = smem + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    const uint32_t staging_base = bar_base + 256u; // 16-byte aligned: bar_base is 1024-aligned
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8); // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // split == 2 ("grouped", K-major split buffers, BN_ <= 128): stages of 3 A pieces + 3 B pieces, kbp k-blocks per tile
    constexpr int G_STAGE = 3 * A_BYTES + 3 * BN_ * BK * 2;
    constexpr int G_STAGES = (STAGES * STAGE_BYTES) / G_STAGE >= 2 ? 2 : 1;
    const bool grouped = p.split == 2 && !A_MN && !B_MN && BN_ <= 128;
    const int num_kb = grouped ? p.kbp : p.split ? 6 * p.kbp : (int)((p.K + BK - 1) / BK);
    const int64_t tiles_per_batch = (int64_t)p.tiles_m * p.tiles_n;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                const int64_t b = t / tiles_per_batch;
                const int r = (int)(t - b * tiles_per_batch);
                const int m0 = (r / p.tiles_n) * BM, n0 = (r % p.tiles_n) * BN_;
                const int ba = (int)b * p.a_bmult, bb = (int)b * p.b_bmult;
                if (grouped) {
                    // fp32 split, narrow tiles: ONE stage carries all three pieces of A and of B for a k-block and the MMA
                    // warp runs the six piece combinations out of it — every operand byte enters shared memory once per
                    // k-block instead of once per combination (PV of attention, N = 64: the A pieces are the traffic)
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        mbar_expect_tx(full_bar(stage), G_STAGE);
                        const uint32_t sa = smem + stage * G_STAGE, sb = sa + 3 * A_BYTES;
                        if (p.conv_taps) { // implicit-GEMM convolution: k-block = (tap, channel block)
                            const int tap = kb / p.conv_cblocks, c0 = (kb - tap * p.conv_cblocks) * BK;
                            const int img = (int)(b / p.conv_oh), oh = (int)(b - (int64_t)img * p.conv_oh) * p.conv_rows;
#pragma unroll
                            for (int pc = 0; pc < 3; ++pc) {
                                tma_load_3d(sa + pc * A_BYTES, &tma_a, full_bar(stage), pc * (int)p.Kp + c0, m0, tap);
                                if (p.conv_rows == 1) {
                                    tma_load_4d(sb + pc * (BN_ * BK * 2), &tma_b, full_bar(stage), pc * (int)p.Kp + c0, n0 + p.conv_qw[tap],
                                                oh + p.conv_qh[tap], img * p.conv_phases + p.conv_ph[tap]);
                                } else { // one 64-column box per output row of the group
                                    for (int r = 0; r < BN_ / 64; ++r)
                                        tma_load_4d(sb + pc * (BN_ * BK * 2) + r * (64 * BK * 2), &tma_b, full_bar(stage), pc * (int)p.Kp + c0,
                                                    p.conv_qw[tap], oh + r * p.conv_rpc + p.conv_qh[tap], img * p.conv_phases + p.conv_ph[tap]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int pc = 0; pc < 3; ++pc) {
                                tma_load_3d(sa + pc * A_BYTES, &tma_a, full_bar(stage), pc * (int)p.Kp + kb * BK, m0, ba);
                                tma_load_3d(sb + pc * (BN_ * BK * 2), &tma_b, full_bar(stage), pc * (int)p.Kp + kb * BK, n0, bb);
                            }
                        }
                        if (++stage == G_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    mbar_expect_tx(full_bar(stage), A_BYTES + BN_ * BK * 2);
                    const uint32_t sa = smem + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    int k0 = kb * BK, k0b = k0;
                    if (p.split) { // K coordinate inside the piece each operand contributes to this combination
                        const int c = kb / p.kbp, kk = (kb - c * p.kbp) * BK;
                        k0 = split_piece_a(c) * (int)p.Kp + kk;
                        k0b = split_piece_b(c) * (int)p.Kp + kk;
                    }
                    if (!A_MN && B_MN && p.conv_taps) {
                        // implicit-GEMM convolution straight from the NCHW tensor (half precision, stride 1): B is MN-major —
                        // a box of 64 columns x 64 channels of input row oh + q_h is exactly one 64-wide N chunk
                        const int tap = kb / p.conv_cblocks, c0 = (kb - tap * p.conv_cblocks) * BK;
                        const int img = (int)(b / p.conv_oh), oh = (int)(b - (int64_t)img * p.conv_oh) * p.conv_rows;
                        tma_load_3d(sa, &tma_a, full_bar(stage), c0, m0, tap);
#pragma unroll
                        for (int j = 0; j < BN_ / 64; ++j) {
                            if (p.conv_rows == 1)
                                tma_load_4d(sb + j * (64 * BK * 2), &tma_b, full_bar(stage), n0 + 64 * j + p.conv_qw[tap], oh + p.conv_qh[tap], c0, img);
                            else
                                tma_load_4d(sb + j * (64 * BK * 2), &tma_b, full_bar(stage), p.conv_qw[tap], oh + j + p.conv_qh[tap], c0, img);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    if (!A_MN && !B_MN && p.conv_taps) { // implicit-GEMM convolution, half-precision operands
                        const int tap = kb / p.conv_cblocks, c0 = (kb - tap * p.conv_cblocks) * BK;
                        const int img = (int)(b / p.conv_oh), oh = (int)(b - (int64_t)img * p.conv_oh) * p.conv_rows;
                        tma_load_3d(sa, &tma_a, full_bar(stage), c0, m0, tap);
                        if (p.conv_rows == 1) {
                            tma_load_4d(sb, &tma_b, full_bar(stage), c0, n0 + p.conv_qw[tap], oh + p.conv_qh[tap],
                                        img * p.conv_phases + p.conv_ph[tap]);
                        } else { // one 64-column box per output row of the group
                            for (int r = 0; r < BN_ / 64; ++r)
                                tma_load_4d(sb + r * (64 * BK * 2), &tma_b, full_bar(stage), c0, p.conv_qw[tap], oh + r * p.conv_rpc + p.conv_qh[tap],
                                            img * p.conv_phases + p.conv_ph[tap]);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    if (!A_MN) {
                        tma_load_3d(sa, &tma_a, full_bar(stage), k0, m0, ba); // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) // box {64 m, 64 k} per 64-wide M chunk
                            tma_load_3d(sa + j * (64 * BK * 2), &tma_a, full_bar(stage), m0 + 64 * j, k0, ba);
                    }
                    if (!B_MN) {
                        tma_load_3d(sb, &tma_b, full_bar(stage), k0b, n0, bb); // box {64 k, 256 n}
                    } else {
#pragma unroll
                        for (int j = 0; j < BN_ / 64; ++j)
                            tma_load_3d(sb + j * (64 * BK * 2), &tma_b, full_bar(stage), n0 + 64 * j, k0b, bb);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        { // whole warp, one elected lane issues (see elect_one)
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u); // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;  // accumulators sit 256 columns apart for every BN_
                if (grouped) {
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t sa = smem + stage * G_STAGE, sb = sa + 3 * A_BYTES;
                        // descriptors: base + (byte offset >> 4) in the start-address field (shared memory < 256 KiB)
                        const uint64_t da = umma_desc(sa, 16, 1024), db = umma_desc(sb, 16, 1024);
                        if (elect_one()) {
#pragma unroll
                            for (int c = 0; c < 6; ++c) {
                                // hi.hi (c == 5) accumulates in the tile's main accumulator; the five small products go to a
                                // SECOND accumulator in the unused upper half of the tile's 256 TMEM columns and meet the big
                                // sum only once, in the epilogue: interleaving them per k-block into one fp32 accumulator
                                // lost their low bits against the running hi.hi sum (3.9e-6 |A||B| on the wide-range test).
                                const uint64_t pa = da + (uint64_t)((split_piece_a(c) * A_BYTES) >> 4);
                                const uint64_t pb = db + (uint64_t)((split_piece_b(c) * (BN_ * BK * 2)) >> 4);
                                const uint32_t dst = c == 5 ? tmem_d : tmem_d + 128u;
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    tc_mma(dst, pa + (uint64_t)(k * 2), pb + (uint64_t)(k * 2), p.idesc,
                                           c == 5 ? (kb | k) != 0 : (kb | c | k) != 0);
                            }
                            tc_commit(empty_bar(stage));
                            if (kb == num_kb - 1)
                                tc_commit(tfull_bar(acc));
                        }
                        __syncwarp();
                        if (++stage == G_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1u;
                    }
                    continue;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    // K-major SW128: rows of 128 B, 8-row groups 1024 B apart (SBO); a K step of 16 elements = +32 B.
                    // MN-major SW128: 64-element MN chunks 8 KiB apart (LBO), 8-row K groups 1024 B apart (SBO);
                    //                 a K step of 16 rows = +2048 B.
                    const uint64_t da = A_MN ? umma_desc(sa, 64 * BK * 2, 1024) : umma_desc(sa, 16, 1024);
                    const uint64_t db = B_MN ? umma_desc(sb, 64 * BK * 2, 1024) : umma_desc(sb, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma(tmem_d, da + (uint64_t)(k * (A_MN ? 128 : 2)), db + (uint64_t)(k * (B_MN ? 128 : 2)), p.idesc, (kb | k) != 0);
                        tc_commit(empty_bar(stage)); // slot is free once these MMAs have read it
                        if (kb == num_kb - 1)
                            tc_commit(tfull_bar(acc)); // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // epilogue: EIGHT warps. Warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); warps 2..5 and 6..9 cover the
        // four lane quarters twice and take alternate 32-column chunks, so every SM sub-partition runs two epilogue warps.
        // (The first version had four warps, per-element bias loads / dtype branches / software bf16 rounding: ~240
        // instructions per 32 outputs and ~25 us per 128x256 tile — as long as the MMAs of a K = 4096 tile, and 10x the
        // MMAs of a K = 384 Linear layer. ncu: tensor pipe 9-20 % active on whisper's layers.)
        const int q = warp & 3, half = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const int64_t b = t / tiles_per_batch;
            const int r = (int)(t - b * tiles_per_batch);
            const int m0 = (r / p.tiles_n) * BM, n0 = (r % p.tiles_n) * BN_;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * BN;  // accumulators sit 256 columns apart for every BN_
            epilogue_tile<BN_>(p, b, m0, n0, taddr, grouped, q, half, lane, warp - 2, staging_base);
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2 --------------------------------------------------------------------
// What bounds the one-CTA kernel on C4 is the operand bytes every SM has to take in per flop: 48 KiB per
// 128x256x64 k-block. Here a CLUSTER OF TWO CTAs (an SM pair) owns one 256x256 tile: each CTA loads its own 128 rows of A
// and only HALF of the B tile (128 of the 256 columns), the leader CTA issues one M = 256 MMA that reads both halves out
// of both shared memories, and each CTA ends up with its 128 rows of the accumulator in its own TMEM. 32 KiB per k-block
// and SM instead of 48, which also leaves room for a 6-deep ring.
//   producer (both CTAs)  : waits on its OWN empty barrier, signals the LEADER's full barrier (TMA .cta_group::2)
//   MMA issuer (leader)   : tcgen05.commit ... multicast::cluster releases the ring slot / publishes the accumulator in
//                           BOTH CTAs with one instruction
//   epilogue (both CTAs)  : same code as the one-CTA kernel on the CTA's 128 rows; arrives remotely on the leader's
//                           "accumulator drained" barrier (count = 16 warps)
constexpr int P_STAGES = 6;
constexpr int PB_BYTES = 128 * BK * 2;            // this CTA's half of the 256-wide B tile
constexpr int P_STAGE_BYTES = A_BYTES + PB_BYTES; // 32 KiB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256 + STAGING_BYTES;
static_assert(P_SMEM_BYTES <= 232448, "CTA-pair GEMM: shared memory");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_shared_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}
// destination = this CTA's shared memory, completion = the barrier at `cluster_bar` (the leader's)
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, uint32_t cluster_bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) { // arrives on the barrier at this offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}

// Tile order inside one matrix: bands of GROUP_M tile rows, walked column by column, so the 74 tiles in flight cover about
// 8 x 9 tiles (A: 2048 rows, B: 2304 columns) instead of 2.3 full tile rows — operands larger than L2 (8192^3: B alone
// is 128 MiB) are then re-read from HBM once per band instead of once per 2.3 tile rows.
constexpr int GROUP_M = 8;
__device__ __forceinline__ void pair_tile_coords(int r, int tiles_m, int tiles_n, int group_m, int &tm, int &tn) {
    const int band = r / (group_m * tiles_n);
    const int first_m = band * group_m;
    const int rows = min(group_m, tiles_m - first_m);
    const int in_band = r - band * group_m * tiles_n;
    tm = first_m + in_band % rows;
    tn = in_band / rows;
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem = (raw + 1023u) & ~1023u;
    const uint32_t bar_base = smem + P_STAGES * P_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * P_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * P_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * P_STAGES + 4);
    const uint32_t staging_base = bar_base + 256u;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank(); // 0 = leader
    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(full_bar(s), 1);  // the leader's arrive.expect_tx (covers the bytes of both CTAs)
            mbar_init(empty_bar(s), 1); // one multicast commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 16); // eight epilogue warps in each of the two CTAs (leader's copy is the one used)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) { // the same warp of both CTAs: the allocation is collective over the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();    // (the CTA-level barrier is what compute-sanitizer's racecheck tracks for the TMEM-slot read below)
    cluster_sync_all(); // barriers of BOTH CTAs exist before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const int num_kb = (int)((p.K + BK - 1) / BK);
    const int64_t tiles_per_batch = (int64_t)p.tiles_m * p.tiles_n;
    const int64_t first = blockIdx.x >> 1, step = gridDim.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = first; t < p.num_tiles; t += step) {
                const int64_t b = t / tiles_per_batch;
                const int r = (int)(t - b * tiles_per_batch);
                int tm, tn;
                pair_tile_coords(r, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
                const int m0 = tm * 256 + (int)rank * 128, n0 = tn * 256 + (int)rank * 128;
                const int ba = (int)b * p.a_bmult, bb = (int)b * p.b_bmult;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (rank == 0)
                        mbar_expect_tx(full_bar(stage), 2 * P_STAGE_BYTES);
                    const uint32_t lfull = map_shared_rank(full_bar(stage), 0);
                    const uint32_t sa = smem + stage * P_STAGE_BYTES, sb = sa + A_BYTES;
                    const int k0 = kb * BK;
                    if (!A_MN) {
                        tma_load_3d_pair(sa, &tma_a, lfull, k0, m0, ba); // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j) // box {64 m, 64 k}
                            tma_load_3d_pair(sa + j * (64 * BK * 2), &tma_a, lfull, m0 + 64 * j, k0, ba);
                    }
                    if (!B_MN) {
                        tma_load_3d_pair(sb, &tma_b, lfull, k0, n0, bb); // box {64 k, 128 n}
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            tma_load_3d_pair(sb + j * (64 * BK * 2), &tma_b, lfull, n0 + 64 * j, k0, bb);
                    }
                    if (++stage == P_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) { // whole warp of the leader CTA, one elected lane issues (see elect_one)
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int64_t t = first; t < p.num_tiles; t += step) {
                mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u); // both epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem + stage * P_STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t da = A_MN ? umma_desc(sa, 64 * BK * 2, 1024) : umma_desc(sa, 16, 1024);
                    const uint64_t db = B_MN ? umma_desc(sb, 64 * BK * 2, 1024) : umma_desc(sb, 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma_pair(tmem_d, da + (uint64_t)(k * (A_MN ? 128 : 2)), db + (uint64_t)(k * (B_MN ? 128 : 2)), p.idesc, (kb | k) != 0);
                        tc_commit_pair(empty_bar(stage));
                        if (kb == num_kb - 1)
                            tc_commit_pair(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == P_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t t = first; t < p.num_tiles; t += step) {
            const int64_t b = t / tiles_per_batch;
            const int r = (int)(t - b * tiles_per_batch);
            int tm, tn;
            pair_tile_coords(r, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
            const int m0 = tm * 256 + (int)rank * 128, n0 = tn * 256;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
            epilogue_tile<256>(p, b, m0, n0, taddr, false, q, half, lane, warp - 2, staging_base);
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                mbar_arrive_remote(map_shared_rank(tempty_bar(acc), 0));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    cluster_sync_all(); // the peer's shared memory and TMEM stay alive until the leader's last MMA and commit have landed
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// collapses the batch dims of one operand to a single (extent, byte stride); false if impossible
bool collapse_batch(const Desc &d, const int64_t *bs, int64_t *stride, int *mult) {
    bool all_zero = true;
    for (int i = 0; i < d.nb; ++i)
        if (d.bshape[i] != 1 && bs[i] != 0)
            all_zero = false;
    if (d.batch == 1 || all_zero) {
        *stride = 0;
        *mult = 0;
        return true;
    }
    int64_t expect = -1, inner = 0;
    for (int i = 0; i < d.nb; ++i) { // innermost first
        if (d.bshape[i] == 1)
            continue;
        if (expect < 0) {
            inner = bs[i];
            expect = bs[i] * d.bshape[i];
        } else {
            if (bs[i] != expect)
                return false;
            expect *= d.bshape[i];
        }
    }
    *stride = inner;
    *mult = 1;
    return inner > 0;
}

struct Operand {
    bool mn_major;
    int64_t ld; // byte stride of the non-contiguous matrix dim
};
// rows/cols strides of the logical [MN][K] operand -> major-ness; false if neither dim is contiguous
bool classify(int64_t mn_stride, int64_t k_stride, int64_t mn_extent, int64_t k_extent, Operand *o) {
    if (k_stride == 2 || k_extent == 1) { // contraction dim contiguous -> K-major
        o->mn_major = false;
        o->ld = mn_extent == 1 ? 0 : mn_stride;
        return o->ld >= 0;
    }
    if (mn_stride == 2 || mn_extent == 1) {
        o->mn_major = true;
        o->ld = k_stride;
        return o->ld >= 0;
    }
    return false;
}

axb_status make_map(CUtensorMap *map, const drv::Api *api, int dtype, const void *base, uint64_t d0, uint64_t d1, uint64_t d2,
                    uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1) {
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstr[2] = {s1, s2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = api->TensorMapEncodeTiled(map, dtype == AXB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                                           const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "cuTensorMapEncodeTiled failed: %s", drv::err_str(r));
    return AXB_OK;
}


// ---- fp32 matmul on the bf16 tensor pipe: exact 3-way operand split -------------------------------------------------
// An fp32 value is the exact sum of three bf16 values: hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)
// (8 significand bits each, same exponent range). A.B then needs the six products whose weight is >= 2^-24 of the
// result: hi.hi, hi.mid, mid.hi, hi.lo, lo.hi, mid.mid — every one an ordinary bf16 MMA accumulated in the SAME fp32 TMEM
// accumulator. Dropped terms (mid.lo, lo.mid, lo.lo) are <= 3 * 2^-24 relative, i.e. the result is as accurate as an fp32
// FMA chain (the reference's sgemm), at 6 bf16 MMAs per fp32 MAC: ~1/6 of the bf16 tensor peak instead of the CUDA-core
// rate. tcgen05 has no IEEE-fp32 MMA; kind::tf32 truncates operands to 10 bits and would need the same 3-term split at
// half the bf16 rate, so bf16 x 6 costs the same and reuses the bf16 data path unchanged.
struct SplitParams {
    const char *src;
    uint16_t *dst;
    int64_t rows, K, Kp;
    int64_t rs, cs; // bytes: element (r, k) of the logical [rows][K] operand
    int32_t nb;
    int64_t bshape[kMaxBatchDims], bs[kMaxBatchDims];
};

__device__ __forceinline__ void split3(float x, uint16_t &hi, uint16_t &mid, uint16_t &lo) {
    hi = red::f2bf(x);
    const float r1 = x - red::bf2f(hi); // exact
    mid = red::f2bf(r1);
    const float r2 = r1 - red::bf2f(mid); // exact
    lo = red::f2bf(r2);
}

// ROWS_CONTIG = false (the common case: K is the unit-stride dim): grid (ceil(Kp / 256), ceil(rows / 8), batch), block
// (32, 8): a thread converts 8 consecutive k of one row — two 128-bit loads, three 128-bit stores.
// ROWS_CONTIG = true (a transposed view: rows are unit-stride): grid (Kp / 64, ceil(rows / 32), batch): the 32x64 tile is
// read along rows and turned in shared memory so that writes stay k-contiguous.
template <bool ROWS_CONTIG> __global__ void __launch_bounds__(256) split_f32_kernel(const __grid_constant__ SplitParams p) {
    __shared__ float tile[ROWS_CONTIG ? 64 : 1][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    int64_t boff = 0, rem = blockIdx.z;
#pragma unroll 1
    for (int i = 0; i < p.nb; ++i) {
        const int64_t q = rem / p.bshape[i];
        boff += (rem - q * p.bshape[i]) * p.bs[i];
        rem = q;
    }
    const char *src = p.src + boff;
    uint16_t *dst = p.dst + (int64_t)blockIdx.z * p.rows * 3 * p.Kp;
    if (!ROWS_CONTIG) {
        const int64_t r = (int64_t)blockIdx.y * 8 + ty, k = ((int64_t)blockIdx.x * 32 + tx) * 8;
        if (r >= p.rows || k >= p.Kp)
            return;
        float v[8];
        const char *rp = src + r * p.rs;
        if (p.cs == 4 && k + 8 <= p.K && ((reinterpret_cast<uintptr_t>(rp) + k * 4) & 15) == 0) {
            const float4 a = *reinterpret_cast<const float4 *>(rp + k * 4), b = *reinterpret_cast<const float4 *>(rp + k * 4 + 16);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = k + i < p.K ? *reinterpret_cast<const float *>(rp + (k + i) * p.cs) : 0.0f;
        }
        uint32_t h[4], m[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            red::split3_pair(v[2 * i], v[2 * i + 1], h[i], m[i], l[i]);
        uint16_t *row = dst + r * 3 * p.Kp + k; // k % 8 == 0 and Kp % 64 == 0: 16-byte aligned
        *reinterpret_cast<uint4 *>(row) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(row + p.Kp) = make_uint4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<uint4 *>(row + 2 * p.Kp) = make_uint4(l[0], l[1], l[2], l[3]);
        return;
    }
    const int64_t k0 = (int64_t)blockIdx.x * 64, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) { // k = k0 + ty + 8 i, r = r0 + tx
        const int64_t k = k0 + ty + 8 * i, r = r0 + tx;
        tile[ROWS_CONTIG ? ty + 8 * i : 0][tx] = (k < p.K && r < p.rows) ? *reinterpret_cast<const float *>(src + r * p.rs + k * p.cs) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        if (r >= p.rows)
            continue;
        const int64_t k = k0 + 2 * tx;
        const float v0 = tile[ROWS_CONTIG ? 2 * tx : 0][ty + 8 * i], v1 = tile[ROWS_CONTIG ? 2 * tx + 1 : 0][ty + 8 * i];
        uint16_t h0, m0, l0, h1, m1, l1;
        split3(v0, h0, m0, l0);
        split3(v1, h1, m1, l1);
        uint16_t *row = dst + r * 3 * p.Kp + k; // k even, Kp % 64 == 0: 4-byte aligned
        *reinterpret_cast<uint32_t *>(row) = (uint32_t)h0 | ((uint32_t)h1 << 16);
        *reinterpret_cast<uint32_t *>(row + p.Kp) = (uint32_t)m0 | ((uint32_t)m1 << 16);
        *reinterpret_cast<uint32_t *>(row + 2 * p.Kp) = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
}

axb_status run_split(const Desc &d, const char *src, int64_t rows, int64_t rs, int64_t cs, const int64_t *bs, bool batched,
                     int64_t Kp, uint16_t *dst, cudaStream_t s) {
    SplitParams sp;
    memset(&sp, 0, sizeof sp);
    sp.src = src; sp.dst = dst; sp.rows = rows; sp.K = d.K; sp.Kp = Kp; sp.rs = rs; sp.cs = cs;
    sp.nb = batched ? d.nb : 0;
    for (int i = 0; i < sp.nb; ++i) {
        sp.bshape[i] = d.bshape[i];
        sp.bs[i] = bs[i];
    }
    const int64_t nbatch = batched ? d.batch : 1;
    const bool rows_contig = rs == 4 && cs != 4;
    const int64_t gy = rows_contig ? (rows + 31) / 32 : (rows + 7) / 8;
    AXB_REQUIRE(nbatch <= 65535 && gy <= 65535, AXB_ERR_UNSUPPORTED, "fp32 tensor-core matmul: operand too large to split");
    dim3 block(32, 8);
    if (rows_contig)
        split_f32_kernel<true><<<dim3((unsigned)(Kp / 64), (unsigned)gy, (unsigned)nbatch), block, 0, s>>>(sp);
    else
        split_f32_kernel<false><<<dim3((unsigned)((Kp + 255) / 256), (unsigned)gy, (unsigned)nbatch), block, 0, s>>>(sp);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

axb_status launch_ff(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int sm_count, cudaStream_t s);
int pick_bn(int64_t N);

// ---- split cache: the pieces of a CONSTANT fp32 operand are computed once ----------------------------------------------
// whisper's encoder re-split every Linear / conv weight on every forward (80 of its 164 launches, 12.6 % of the GPU time,
// profiles/r1_launches_whisper_b8_s6.csv). An operand flagged AXB_MM_CONST_* keeps its [rows][3 * Kp] pieces in a device
// buffer keyed by (address, layout); the entry dies when its source bytes are freed or overwritten.
struct SplitKey {
    const char *src;
    int64_t rows, K, rs, cs, batch;
    int64_t bs[kMaxBatchDims], bshape[kMaxBatchDims];
    int32_t nb, device;
    bool operator==(const SplitKey &o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
struct SplitEntry {
    SplitKey key;
    const char *lo, *hi; // source byte range
    void *pieces;
    size_t bytes;
    uint64_t tick;
};
std::mutex g_split_mu;
std::vector<SplitEntry> g_split;
std::atomic<int> g_split_n{0};
axb_const_cache_stats g_split_stats{};
uint64_t g_split_tick = 0;
constexpr size_t kSplitCacheBytes = 4ull << 30;

SplitKey split_key(const Desc &d, const char *src, int64_t rows, int64_t rs, int64_t cs, const int64_t *bs, bool batched, int device) {
    SplitKey k;
    memset(&k, 0, sizeof k);
    k.src = src; k.rows = rows; k.K = d.K; k.rs = rs; k.cs = cs; k.device = device;
    k.batch = batched ? d.batch : 1;
    k.nb = batched ? d.nb : 0;
    for (int i = 0; i < k.nb; ++i) {
        k.bs[i] = bs[i];
        k.bshape[i] = d.bshape[i];
    }
    return k;
}
// returns the cached pieces or nullptr
void *split_cache_find(const SplitKey &k) {
    if (g_split_n.load(std::memory_order_acquire) == 0)
        return nullptr;
    std::lock_guard<std::mutex> lk(g_split_mu);
    for (auto &e : g_split)
        if (e.key == k) {
            e.tick = ++g_split_tick;
            g_split_stats.hits++;
            return e.pieces;
        }
    return nullptr;
}
void split_cache_insert(const SplitKey &k, void *pieces, size_t bytes, cudaStream_t s) {
    // byte range the operand covers (non-negative strides only; others are never cached)
    int64_t span = 4 + (k.rows - 1) * k.rs + (k.K - 1) * k.cs;
    for (int i = 0; i < k.nb; ++i)
        span += (k.bshape[i] - 1) * k.bs[i];
    std::vector<void *> evict;
    {
        std::lock_guard<std::mutex> lk(g_split_mu);
        g_split_stats.misses++;
        g_split.push_back(SplitEntry{k, k.src, k.src + span, pieces, bytes, ++g_split_tick});
        g_split_stats.bytes += bytes;
        while (g_split_stats.bytes > kSplitCacheBytes && g_split.size() > 1) { // least recently used first
            size_t v = 0;
            for (size_t i = 1; i + 1 < g_split.size(); ++i)
                if (g_split[i].tick < g_split[v].tick)
                    v = i;
            evict.push_back(g_split[v].pieces);
            g_split_stats.bytes -= g_split[v].bytes;
            g_split.erase(g_split.begin() + v);
        }
        g_split_stats.entries = g_split.size();
        g_split_n.store((int)g_split.size(), std::memory_order_release);
    }
    for (void *q : evict)
        axb_free(q, s);
}

axb_status try_split_f32(const Desc &d, DeviceCtx *ctx, cudaStream_t s, bool *done) {
    const PresplitA pre = presplit_a_hint();
    presplit_a_hint() = PresplitA{};
    static const bool disabled = getenv("AXB_F32_MATMUL_SIMT") != nullptr; // A/B switch
    if (disabled && !pre.ptr)
        return AXB_OK;
    const double flops = 2.0 * (double)d.M * (double)d.N * (double)d.K * (double)d.batch;
    const bool small = flops < 1.0e8 || d.M < 64 || d.N < 64 || d.K < 32;
    if ((small && !pre.ptr) || d.M >= (1ll << 31) || d.N >= (1ll << 31) || d.K >= (1ll << 28) || d.batch > 65535) {
        AXB_REQUIRE(!pre.ptr, AXB_ERR_UNSUPPORTED, "fp32 matmul with a pre-split operand: problem outside the tensor-core path");
        return AXB_OK; // small problems: the split passes and the 128x256 tile would dominate
    }
    int64_t c_bs, bias_bs = 0;
    int c_mult, bias_mult = 0;
    if (!collapse_batch(d, d.c_bs, &c_bs, &c_mult) || (d.batch > 1 && !c_mult) ||
        (d.bias && !collapse_batch(d, d.bias_bs, &bias_bs, &bias_mult)) || d.c_rs < 0 || d.c_cs < 0) {
        AXB_REQUIRE(!pre.ptr, AXB_ERR_UNSUPPORTED, "fp32 matmul with a pre-split operand: output layout outside the tensor-core path");
        return AXB_OK;
    }
    auto batched = [&](const int64_t *bs) {
        for (int i = 0; i < d.nb; ++i)
            if (d.bshape[i] != 1 && bs[i] != 0)
                return true;
        return false;
    };
    const bool a_b = d.batch > 1 && batched(d.a_bs), b_b = d.batch > 1 && batched(d.b_bs);
    const int64_t Kp = (d.K + 63) / 64 * 64;
    AXB_REQUIRE(!pre.ptr || pre.kp == Kp, AXB_ERR_INVALID, "fp32 matmul with a pre-split operand: piece width %lld, expected %lld",
                (long long)pre.kp, (long long)Kp);
    const size_t a_bytes = (size_t)(a_b ? d.batch : 1) * d.M * 3 * Kp * 2, b_bytes = (size_t)(b_b ? d.batch : 1) * d.N * 3 * Kp * 2;
    const drv::Api *api = drv::api();
    if (!api)
        return AXB_ERR_CUDA;
    void *wa = nullptr, *wb = nullptr;
    // x.matmul(x.T()) (the README chain): B viewed as [N][K] is the very same strided matrix as A — one split serves both
    bool same_operand = !pre.ptr && d.a == d.b && d.M == d.N && d.a_rs == d.b_cs && d.a_cs == d.b_rs && a_b == b_b;
    for (int i = 0; i < d.nb && same_operand; ++i)
        same_operand = d.a_bs[i] == d.b_bs[i];
    // constant operands (weights): pieces from the split cache; a miss computes them once and keeps them
    auto cacheable = [&](int flag, int64_t rs, int64_t cs, const int64_t *bs, bool batched) {
        if (!(d.const_ops & flag) || rs < 0 || cs < 0)
            return false;
        for (int i = 0; batched && i < d.nb; ++i)
            if (bs[i] < 0)
                return false;
        return true;
    };
    const bool a_const = !pre.ptr && cacheable(AXB_MM_CONST_A, d.a_rs, d.a_cs, d.a_bs, a_b);
    const bool b_const = !same_operand && cacheable(AXB_MM_CONST_B, d.b_cs, d.b_rs, d.b_bs, b_b);
    SplitKey ka{}, kb{};
    bool a_hit = false, b_hit = false;
    if (a_const) {
        ka = split_key(d, d.a, d.M, d.a_rs, d.a_cs, d.a_bs, a_b, ctx->device);
        wa = split_cache_find(ka);
        a_hit = wa != nullptr;
    }
    if (b_const) {
        kb = split_key(d, d.b, d.N, d.b_cs, d.b_rs, d.b_bs, b_b, ctx->device);
        wb = split_cache_find(kb);
        b_hit = wb != nullptr;
    }
    if (pre.ptr)
        wa = const_cast<uint16_t *>(pre.ptr);
    else if (!wa)
        AXB_TRY(axb_alloc(a_bytes, s, &wa));
    axb_status st = AXB_OK;
    if (same_operand)
        wb = wa;
    else if (!wb)
        st = axb_alloc(b_bytes, s, &wb);
    if (st != AXB_OK) {
        if (!pre.ptr && !a_hit)
            axb_free(wa, s);
        return st;
    }
    bool a_keep = a_hit, b_keep = b_hit; // owned by the cache: not returned to the pool here
    auto cleanup = [&]() {
        if (!pre.ptr && !a_keep)
            axb_free(wa, s);
        if (!same_operand && !b_keep)
            axb_free(wb, s);
    };
    // A is [M][K] with strides (a_rs, a_cs); B is used as [N][K] with strides (b_cs, b_rs)
    if (!pre.ptr && !a_hit)
        st = run_split(d, d.a, d.M, d.a_rs, d.a_cs, d.a_bs, a_b, Kp, static_cast<uint16_t *>(wa), s);
    if (st == AXB_OK && !same_operand && !b_hit)
        st = run_split(d, d.b, d.N, d.b_cs, d.b_rs, d.b_bs, b_b, Kp, static_cast<uint16_t *>(wb), s);
    if (st == AXB_OK && a_const && !a_hit) {
        split_cache_insert(ka, wa, a_bytes, s);
        a_keep = true;
    }
    if (st == AXB_OK && b_const && !b_hit) {
        split_cache_insert(kb, wb, b_bytes, s);
        b_keep = true;
    }
    // grouped stages need BN <= 128; they also carry more flops per loaded byte (131 vs 87 flop/B for the streamed 256-wide
    // tile), so the split path prefers 128-wide tiles whenever N > 64
    static const bool wide_split = getenv("AXB_SPLIT_BN256") != nullptr; // A/B switch
    const int bn = wide_split ? pick_bn(d.N) : (d.N <= 64 ? 64 : 128);
    CUtensorMap ma, mb;
    if (st == AXB_OK)
        st = make_map(&ma, api, AXB_BF16, wa, (uint64_t)(3 * Kp), (uint64_t)d.M, a_b ? (uint64_t)d.batch : 1, (uint64_t)(3 * Kp * 2),
                      (uint64_t)d.M * 3 * Kp * 2, BK, BM);
    if (st == AXB_OK)
        st = make_map(&mb, api, AXB_BF16, wb, (uint64_t)(3 * Kp), (uint64_t)d.N, b_b ? (uint64_t)d.batch : 1, (uint64_t)(3 * Kp * 2),
                      (uint64_t)d.N * 3 * Kp * 2, BK, (uint32_t)bn);
    if (st != AXB_OK) {
        cleanup();
        return st;
    }
    TcParams p;
    memset(&p, 0, sizeof p);
    p.M = d.M; p.N = d.N; p.K = d.K; p.batch = d.batch;
    p.tiles_m = (int)((d.M + BM - 1) / BM);
    p.bn = bn;
    p.tiles_n = (int)((d.N + bn - 1) / bn);
    p.num_tiles = (int64_t)p.tiles_m * p.tiles_n * d.batch;
    p.c = d.c; p.c_rs = d.c_rs; p.c_cs = d.c_cs; p.c_bs = c_bs;
    p.bias = d.bias; p.bias_rs = d.bias_rs; p.bias_cs = d.bias_cs; p.bias_bs = bias_mult ? bias_bs : 0;
    p.bias_dtype = d.bias_dtype; p.act = d.act; p.out_dtype = AXB_F32;
    p.a_bmult = a_b ? 1 : 0; p.b_bmult = b_b ? 1 : 0;
    static const bool no_group = getenv("AXB_SPLIT_NO_GROUP") != nullptr; // A/B switch
    p.split = (bn <= 128 && !no_group) ? 2 : 1; p.kbp = (int)(Kp / BK); p.Kp = Kp;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    st = launch_ff(ma, mb, p, ctx->sm_count, s);
    cleanup(); // stream-ordered: the blocks return to the pool after the GEMM in stream order
    if (st == AXB_OK) {
        *done = true;
        last_path() = 3;
    }
    return st;
}

template <bool A_MN, bool B_MN, int BN_>
axb_status launch_bn(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int sm_count, cudaStream_t s) {
    static bool configured = false; // per instantiation; benign race (same value)
    if (!configured) {
        AXB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<A_MN, B_MN, BN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, sm_count);
    gemm_tc_kernel<A_MN, B_MN, BN_><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(ma, mb, p);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}
template <bool A_MN, bool B_MN>
axb_status launch(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int sm_count, cudaStream_t s) {
    switch (p.bn) {
    case 64: return launch_bn<A_MN, B_MN, 64>(ma, mb, p, sm_count, s);
    case 128: return launch_bn<A_MN, B_MN, 128>(ma, mb, p, sm_count, s);
    default: return launch_bn<A_MN, B_MN, 256>(ma, mb, p, sm_count, s);
    }
}
// tile width: the narrowest of 64 / 128 / 256 that wastes the least of the last tile (ties -> wider)
int pick_bn(int64_t N) {
    static const char *force = getenv("AXB_GEMM_BN");
    if (force)
        return atoi(force) == 64 ? 64 : atoi(force) == 128 ? 128 : 256;
    int best = 256;
    int64_t best_pad = (N + 255) / 256 * 256;
    for (int bn : {128, 64}) {
        const int64_t pad = (N + bn - 1) / bn * bn;
        if (pad < best_pad) {
            best = bn;
            best_pad = pad;
        }
    }
    return best;
}

axb_status launch_ff(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int sm_count, cudaStream_t s) {
    return launch<false, false>(ma, mb, p, sm_count, s);
}

template <bool A_MN, bool B_MN>
axb_status launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, const TcParams &p, int sm_count, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        AXB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
        configured = true;
    }
    const unsigned grid = 2u * (unsigned)std::min<int64_t>(p.num_tiles, sm_count / 2); // one cluster (SM pair) per 256x256 tile
    gemm_tc_pair_kernel<A_MN, B_MN><<<grid, NUM_THREADS, P_SMEM_BYTES, s>>>(ma, mb, p);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}
// The CTA-pair kernel takes every problem whose 256x256 tiles fill the 74 SM pairs at least twice; AXB_GEMM_PAIR=0/1 forces it.
bool use_pair(const Desc &d, int bn, int sm_count) {
    static const char *force = getenv("AXB_GEMM_PAIR");
    if (bn != 256 || d.M <= 128)
        return false;
    if (force)
        return atoi(force) != 0;
    const int64_t pad256 = (d.M + 255) / 256 * 256, pad128 = (d.M + 127) / 128 * 128;
    if ((double)pad256 > 1.04 * (double)pad128)
        return false; // a ragged last tile whose second half would be all padding
    const int64_t tiles = (pad256 / 256) * ((d.N + 255) / 256) * d.batch;
    return tiles >= 2 * (sm_count / 2);
}


// ---- implicit-GEMM convolution --------------------------------------------------------------------------------------------
// out[n, co, oh, ow] = sum over (kh, kw, ci) of w[co, ci, kh, kw] * x[n, ci, oh*sh + kh*dh - ph, ow*sw + kw*dw - pw]
// is, for one output row (n, oh), the GEMM  C[co][ow] = sum over taps t = (kh, kw) of  W_t[co][ci] . X_t[ci][ow]  where X_t is
// the input row shifted by the tap. Nothing is gathered into an im2col matrix: the K loop of gemm_tc_kernel walks (tap,
// 64-channel block) steps and the TMA load of the B tile simply starts at the shifted coordinate — columns / rows that fall
// outside the image come back as zeros (TMA out-of-bounds fill), which is exactly the convolution's zero padding.
// Two preparation passes put the operands into the K-major form the kernel consumes:
//   conv_prep_x   NC(H)W -> channels-last, de-interleaved by stride phase: [image * sh*sw + rh*sw + rw][H'][W'][pieces * Cp]
//                 (h = h'*sh + rh, w = w'*sw + rw), so a strided convolution also reads unit-stride rows; fp32 inputs are
//                 written as their three bf16 pieces (the exact split of try_split_f32), half inputs are transposed as is
//   conv_prep_w   (C_out, C_in, KH, KW) -> [tap][C_out][pieces * Cp]
// Reference being replaced: the scalar loops of ops::conv1d / conv2d (src/tensor/operations.cpp:2968-3330) and the MPSGraph
// lowering (src/backends/metal/mpsgraph_graph_compiler.mm:365-405).
struct ConvPrepX {
    const char *x;
    uint16_t *dst;
    int64_t xs[4];     // byte strides of (N, C, H, W)
    int64_t C, H, W, Cp, Hq, Wq;
    int32_t sh, sw, pieces, src_f32;
    int32_t wtiles;
};
// block (32, 8). A tile of 64 channels x TW columns (TW = 64 for 2-byte inputs, 32 for fp32) is read along w' — a warp reads
// 128 contiguous bytes of one channel row — turned in shared memory and written along channels, 128 bytes per warp and piece
// (the first version moved 32 x 32 tiles with 64-byte rows and ran at 1.1 TB/s: 35 % of a 256-channel 3x3 convolution).
template <bool F32> __global__ void __launch_bounds__(256) conv_prep_x_kernel(const __grid_constant__ ConvPrepX q) {
    constexpr int TW = F32 ? 32 : 64;
    __shared__ uint32_t tile[TW][F32 ? 65 : 33]; // fp32: [w'][c] words; half: [w'][c pair] words (two channels per word)
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int wt = blockIdx.x % q.wtiles, ct = blockIdx.x / q.wtiles;
    const int64_t hq = blockIdx.y, plane = blockIdx.z;
    const int phases = q.sh * q.sw;
    const int64_t n = plane / phases;
    const int ph = (int)(plane - n * phases), rh = ph / q.sw, rw = ph - rh * q.sw;
    const int64_t ih = hq * q.sh + rh;
    const int64_t c0 = (int64_t)ct * 64, w0 = (int64_t)wt * TW;
    const bool row_ok = ih < q.H;
    if (F32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { // c = c0 + ty + 8 i, w' = w0 + tx
            const int64_t c = c0 + ty + 8 * i, wq = w0 + tx, iw = wq * q.sw + rw;
            float v = 0.0f;
            if (row_ok && c < q.C && wq < q.Wq && iw < q.W)
                v = *reinterpret_cast<const float *>(q.x + n * q.xs[0] + c * q.xs[1] + ih * q.xs[2] + iw * q.xs[3]);
            tile[tx][ty + 8 * i] = __float_as_uint(v);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) { // w' = w0 + ty + 8 i, channels c0 + 2 tx, + 1
            const int64_t wq = w0 + ty + 8 * i, c = c0 + 2 * tx;
            if (wq >= q.Wq || c >= q.Cp)
                continue;
            uint16_t h0, m0, l0, h1, m1, l1;
            split3(__uint_as_float(tile[ty + 8 * i][2 * tx]), h0, m0, l0);
            split3(__uint_as_float(tile[ty + 8 * i][2 * tx + 1]), h1, m1, l1);
            uint16_t *row = q.dst + (((plane * q.Hq + hq) * q.Wq + wq) * 3) * q.Cp + c; // c even, Cp % 64 == 0: 4-byte aligned
            *reinterpret_cast<uint32_t *>(row) = (uint32_t)h0 | ((uint32_t)h1 << 16);
            *reinterpret_cast<uint32_t *>(row + q.Cp) = (uint32_t)m0 | ((uint32_t)m1 << 16);
            *reinterpret_cast<uint32_t *>(row + 2 * q.Cp) = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
    } else {
        uint16_t *t16 = reinterpret_cast<uint16_t *>(&tile[0][0]); // element (w', c) at t16[w' * 66 + c]
#pragma unroll
        for (int i = 0; i < 8; ++i) { // c = c0 + ty + 8 i, w' = w0 + 2 tx, + 1
            const int64_t c = c0 + ty + 8 * i, wq = w0 + 2 * tx;
            uint16_t v0 = 0, v1 = 0;
            if (row_ok && c < q.C) {
                const char *rp = q.x + n * q.xs[0] + c * q.xs[1] + ih * q.xs[2];
                const int64_t iw0 = wq * q.sw + rw, iw1 = iw0 + q.sw;
                if (q.sw == 1 && q.xs[3] == 2 && iw1 < q.W && wq + 1 < q.Wq && ((reinterpret_cast<uintptr_t>(rp) + iw0 * 2) & 3) == 0) {
                    const uint32_t two = *reinterpret_cast<const uint32_t *>(rp + iw0 * 2);
                    v0 = (uint16_t)two;
                    v1 = (uint16_t)(two >> 16);
                } else {
                    if (wq < q.Wq && iw0 < q.W)
                        v0 = *reinterpret_cast<const uint16_t *>(rp + iw0 * q.xs[3]);
                    if (wq + 1 < q.Wq && iw1 < q.W)
                        v1 = *reinterpret_cast<const uint16_t *>(rp + iw1 * q.xs[3]);
                }
            }
            t16[(2 * tx) * 66 + ty + 8 * i] = v0;
            t16[(2 * tx + 1) * 66 + ty + 8 * i] = v1;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) { // w' = w0 + ty + 8 i, channels c0 + 2 tx, + 1
            const int64_t wq = w0 + ty + 8 * i, c = c0 + 2 * tx;
            if (wq >= q.Wq || c >= q.Cp)
                continue;
            const uint32_t two = *reinterpret_cast<const uint32_t *>(&t16[(ty + 8 * i) * 66 + 2 * tx]);
            *reinterpret_cast<uint32_t *>(q.dst + ((plane * q.Hq + hq) * q.Wq + wq) * q.Cp + c) = two;
        }
    }
}
struct ConvPrepW {
    const char *w;
    uint16_t *dst;
    int64_t ws[4]; // byte strides of (C_out, C_in, KH, KW)
    int64_t Cout, Cin, Cp, KH, KW;
    int32_t pieces, src_f32;
    int64_t total; // taps * Cout * Cp
};
// grid (Cp / 64, C_out), block 256: the 64 channels x taps block of one output channel is read in memory order (channel-
// major, taps contiguous for a contiguous weight), staged in shared memory and written tap by tap, 64 channels (128 bytes)
// per row and piece
__global__ void __launch_bounds__(256) conv_prep_w_kernel(const __grid_constant__ ConvPrepW q) {
    __shared__ uint32_t stage[64 * 64]; // [c local][tap], raw bits (fp32 word or 16-bit value)
    const int64_t c0 = (int64_t)blockIdx.x * 64, co = blockIdx.y;
    const int taps = (int)(q.KH * q.KW);
    for (int e = threadIdx.x; e < 64 * taps; e += 256) {
        const int cl = e / taps, tap = e - cl * taps;
        const int64_t c = c0 + cl, kh = tap / q.KW, kw = tap - kh * q.KW;
        uint32_t v = 0;
        if (c < q.Cin) {
            const char *sp = q.w + co * q.ws[0] + c * q.ws[1] + kh * q.ws[2] + kw * q.ws[3];
            v = q.src_f32 ? *reinterpret_cast<const uint32_t *>(sp) : (uint32_t) * reinterpret_cast<const uint16_t *>(sp);
        }
        stage[cl * taps + tap] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * taps; e += 256) {
        const int tap = e / 64, cl = e - tap * 64;
        const uint32_t v = stage[cl * taps + tap];
        uint16_t *row = q.dst + ((tap * q.Cout + co) * q.pieces) * q.Cp + c0 + cl;
        if (q.src_f32) {
            uint16_t h, m, l;
            split3(__uint_as_float(v), h, m, l);
            row[0] = h;
            row[q.Cp] = m;
            row[2 * q.Cp] = l;
        } else {
            row[0] = (uint16_t)v;
        }
    }
}

axb_status make_map4(CUtensorMap *map, const drv::Api *api, int dtype, const void *base, const uint64_t dims[4], const uint64_t strides[3],
                     uint32_t b0, uint32_t b1, uint32_t b2 = 1) {
    cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides[0], strides[1], strides[2]};
    cuuint32_t box[4] = {b0, b1, b2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = api->TensorMapEncodeTiled(map, dtype == AXB_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                                           const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed: %s", drv::err_str(r));
    return AXB_OK;
}

} // namespace

axb_status conv_tc(const ConvDesc &d, DeviceCtx *ctx, cudaStream_t s, bool *done) {
    *done = false;
    static const bool off = getenv("AXB_CONV_NO_TC") != nullptr; // A/B switch
    if (off)
        return AXB_OK;
    const bool f32 = d.dtype == AXB_F32;
    if (!(f32 || d.dtype == AXB_BF16 || d.dtype == AXB_F16))
        return AXB_OK;
    const int64_t taps = d.KH * d.KW;
    const double flops = 2.0 * (double)d.N * (double)d.Cout * (double)d.OH * (double)d.OW * (double)d.Cin * (double)taps;
    if (taps < 1 || taps > kMaxTaps || d.N * d.OH >= (1ll << 31))
        return AXB_OK;
    if (!d.force && (flops < 1.0e8 || d.Cout < 32 || d.OW < 8 || d.Cin < 8))
        return AXB_OK; // small problems: the preparation passes and the 128-row tile would dominate (gather kernel / im2col path)
    if (d.sh < 1 || d.sw < 1 || d.sh * d.sw > 64 || d.N * d.sh * d.sw > 65535 || (d.H + d.sh - 1) / d.sh > 65535)
        return AXB_OK;
    const drv::Api *api = drv::api();
    if (!api)
        return AXB_ERR_CUDA;
    const int pieces = f32 ? 3 : 1;
    const int esz = f32 ? 4 : 2;
    const int64_t Cp = (d.Cin + 63) / 64 * 64, cblocks = Cp / 64;
    const int64_t Hq = (d.H + d.sh - 1) / d.sh, Wq = (d.W + d.sw - 1) / d.sw, planes = d.N * d.sh * d.sw;
    int16_t qw[kMaxTaps], qh[kMaxTaps], phs[kMaxTaps];
    for (int64_t kh = 0; kh < d.KH; ++kh)
        for (int64_t kw = 0; kw < d.KW; ++kw) {
            auto floor_split = [](int64_t t, int64_t stride, int64_t *q, int64_t *r) {
                *r = ((t % stride) + stride) % stride;
                *q = (t - *r) / stride;
            };
            int64_t q_h, r_h, q_w, r_w;
            floor_split(kh * d.dh - d.ph, d.sh, &q_h, &r_h);
            floor_split(kw * d.dw - d.pw, d.sw, &q_w, &r_w);
            if (q_h < -30000 || q_h > 30000 || q_w < -30000 || q_w > 30000)
                return AXB_OK;
            const int64_t t = kh * d.KW + kw;
            qh[t] = (int16_t)q_h;
            qw[t] = (int16_t)q_w;
            phs[t] = (int16_t)(r_h * d.sw + r_w);
        }
    // half precision, stride 1, filters ONE column wide without horizontal padding (1x1 / KHx1 convolutions), a tensor TMA can
    // describe as it is (unit-stride columns, 16-byte aligned rows / planes / base): the input is read in place as an MN-major
    // operand — no preparation pass over it at all. Wider filters cannot take this route: a tap would shift the box by single
    // elements along the INNERMOST dimension, and TMA faults on box starts there that are not 16-byte aligned (measured with
    // tools/tma_probe.cu: box {64,1,64,1} at w = -1 -> illegal instruction; shifts in the outer dimensions are fine)
    static const bool no_direct = getenv("AXB_CONV_NO_DIRECT") != nullptr; // A/B switch
    const bool direct = !no_direct && !f32 && d.sh == 1 && d.sw == 1 && d.KW == 1 && d.pw == 0 && d.x_s[3] == 2 && d.x_s[1] > 0 &&
                        d.x_s[1] % 16 == 0 &&
                        (d.N == 1 || (d.x_s[0] > 0 && d.x_s[0] % 16 == 0)) && (d.H == 1 || (d.x_s[2] >= d.W * 2 && d.x_s[2] % 16 == 0)) &&
                        (reinterpret_cast<uintptr_t>(d.x) & 15) == 0;
    const size_t x_bytes = direct ? 256 : (size_t)planes * Hq * Wq * pieces * Cp * 2, w_bytes = (size_t)taps * d.Cout * pieces * Cp * 2;
    void *xq = nullptr, *wq = nullptr;
    const bool x_shared = d.x_prepared && !direct;      // the prepared input outlives this call
    const bool x_ready = x_shared && *d.x_prepared;     // ... and an earlier call already filled it
    if (x_ready)
        xq = *d.x_prepared;
    else
        AXB_TRY(axb_alloc(x_bytes, s, &xq));
    axb_status st = axb_alloc(w_bytes, s, &wq);
    if (st != AXB_OK) {
        if (!x_ready)
            axb_free(xq, s);
        return st;
    }
    auto body = [&]() -> axb_status {
        ConvPrepX px;
        memset(&px, 0, sizeof px);
        if (!direct && !x_ready) {
        px.x = d.x; px.dst = static_cast<uint16_t *>(xq);
        for (int i = 0; i < 4; ++i)
            px.xs[i] = d.x_s[i];
        px.C = d.Cin; px.H = d.H; px.W = d.W; px.Cp = Cp; px.Hq = Hq; px.Wq = Wq;
        px.sh = d.sh; px.sw = d.sw; px.pieces = pieces; px.src_f32 = f32 ? 1 : 0;
        const int tw = f32 ? 32 : 64;
        px.wtiles = (int)((Wq + tw - 1) / tw);
        const int64_t gx = (int64_t)px.wtiles * (Cp / 64);
        AXB_REQUIRE(gx < (1ll << 31), AXB_ERR_UNSUPPORTED, "conv: input too wide for the preparation pass");
        if (f32)
            conv_prep_x_kernel<true><<<dim3((unsigned)gx, (unsigned)Hq, (unsigned)planes), dim3(32, 8), 0, s>>>(px);
        else
            conv_prep_x_kernel<false><<<dim3((unsigned)gx, (unsigned)Hq, (unsigned)planes), dim3(32, 8), 0, s>>>(px);
        AXB_CHECK_LAUNCH();
        }
        ConvPrepW pw;
        memset(&pw, 0, sizeof pw);
        pw.w = d.w; pw.dst = static_cast<uint16_t *>(wq);
        for (int i = 0; i < 4; ++i)
            pw.ws[i] = d.w_s[i];
        pw.Cout = d.Cout; pw.Cin = d.Cin; pw.Cp = Cp; pw.KH = d.KH; pw.KW = d.KW; pw.pieces = pieces; pw.src_f32 = f32 ? 1 : 0;
        pw.total = taps * d.Cout * Cp;
        AXB_REQUIRE(d.Cout <= 65535, AXB_ERR_UNSUPPORTED, "conv: more than 65535 output channels");
        conv_prep_w_kernel<<<dim3((unsigned)(Cp / 64), (unsigned)d.Cout), 256, 0, s>>>(pw);
        AXB_CHECK_LAUNCH();

        const int mdt = d.dtype == AXB_F16 ? AXB_F16 : AXB_BF16;
        int bn = f32 ? (d.OW <= 64 ? 64 : 128) : pick_bn(d.OW);
        // images at most 64 columns wide: a tile holds bn / 64 consecutive output rows (one 64-column chunk each), so the
        // weight tile is fetched once per group of rows instead of once per row (L2 -> SM traffic per flop halves / quarters)
        static const bool no_rows = getenv("AXB_CONV_ONE_ROW") != nullptr; // A/B switch
        int rows_per_tile = 1, rpc = 1;
        if (!no_rows && d.OW <= 64 && d.OH >= 2) {
            static const bool no_pack = getenv("AXB_CONV_NO_PACK") != nullptr; // A/B switch
            // narrow images: two rows of <= 32 columns, or four of <= 16, share a 64-column chunk
            static const bool no_pack4 = getenv("AXB_CONV_NO_PACK4") != nullptr;
            rpc = (no_pack || direct) ? 1 : (!no_pack4 && d.OW <= 16 && d.OH >= 8) ? 4 : (d.OW <= 32 && d.OH >= 4) ? 2 : 1;
            bn = (f32 || d.OH < 4 * rpc) ? 128 : 256;
            if (rpc >= 2) { // the widest tile that still occupies half of the SMs: small problems prefer more, narrower tiles
                const int64_t tm = (d.Cout + BM - 1) / BM; // (measured: 128 wide tiles on 148 SMs beat 256 narrow ones, 64 do not)
                auto rows_of = [&](int w) { return (int64_t)(w / 64 * rpc); };
                while (bn > 64 && tm * d.N * ((d.OH + rows_of(bn) - 1) / rows_of(bn)) < ctx->sm_count / 2)
                    bn >>= 1;
            }
            rows_per_tile = bn / 64 * rpc;
        }
        const int64_t groups = (d.OH + rows_per_tile - 1) / rows_per_tile;
        const uint64_t rowb = (uint64_t)pieces * Cp * 2;
        CUtensorMap ma, mb;
        AXB_TRY(make_map(&ma, api, mdt, wq, (uint64_t)pieces * Cp, (uint64_t)d.Cout, (uint64_t)taps, rowb, (uint64_t)d.Cout * rowb, BK, BM));
        if (direct) { // (w, h, c, n) over the caller's tensor; one box = 64 columns x 64 channels of one input row
            const uint64_t dims[4] = {(uint64_t)d.W, (uint64_t)d.H, (uint64_t)d.Cin, (uint64_t)d.N};
            const uint64_t strides[3] = {(uint64_t)(d.H == 1 ? (d.W * 2 + 15) / 16 * 16 : d.x_s[2]), (uint64_t)d.x_s[1],
                                         (uint64_t)(d.N == 1 ? d.x_s[1] * d.Cin : d.x_s[0])};
            AXB_TRY(make_map4(&mb, api, mdt, d.x, dims, strides, 64, 1, BK));
        } else {
            const uint64_t dims[4] = {(uint64_t)pieces * Cp, (uint64_t)Wq, (uint64_t)Hq, (uint64_t)planes};
            const uint64_t strides[3] = {rowb, (uint64_t)Wq * rowb, (uint64_t)Hq * Wq * rowb};
            AXB_TRY(make_map4(&mb, api, mdt, xq, dims, strides, BK, rows_per_tile > 1 ? (uint32_t)(64 / rpc) : (uint32_t)bn, (uint32_t)rpc));
        }
        TcParams p;
        memset(&p, 0, sizeof p);
        p.M = d.Cout; p.N = d.OW; p.batch = d.N * groups;
        p.K = taps * cblocks * BK; // the plain (half-precision) path derives its k-block count from K
        p.tiles_m = (int)((d.Cout + BM - 1) / BM);
        p.bn = bn;
        p.tiles_n = rows_per_tile > 1 ? 1 : (int)((d.OW + bn - 1) / bn);
        p.num_tiles = (int64_t)p.tiles_m * p.tiles_n * p.batch;
        p.c = d.out;
        p.c_rs = d.OH * d.OW * esz; p.c_cs = esz; p.c_bs = d.Cout * d.OH * d.OW * esz; p.c_hs = d.OW * esz;
        if (d.o_s[0] | d.o_s[1] | d.o_s[2] | d.o_s[3]) { // a strided view of the result (one phase of a transposed convolution)
            p.c_bs = d.o_s[0]; p.c_rs = d.o_s[1]; p.c_hs = d.o_s[2]; p.c_cs = d.o_s[3];
        }
        p.bias = d.bias; p.bias_rs = d.bias ? d.bias_s : 0; p.bias_cs = 0; p.bias_bs = 0;
        p.bias_dtype = d.bias_dtype; p.act = d.act; p.out_dtype = d.dtype;
        p.split = f32 ? 2 : 0; p.kbp = (int)(taps * cblocks); p.Kp = Cp;
        p.conv_taps = (int)taps; p.conv_cblocks = (int)cblocks; p.conv_oh = (int)groups; p.conv_phases = d.sh * d.sw;
        p.conv_rows = rows_per_tile; p.conv_ohs = (int)d.OH; p.conv_rpc = rpc;
        for (int64_t t = 0; t < taps; ++t) {
            p.conv_qw[t] = qw[t];
            p.conv_qh[t] = qh[t];
            p.conv_ph[t] = phs[t];
        }
        const uint32_t fmt = mdt == AXB_BF16 ? 1u : 0u;
        p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((direct ? 1u : 0u) << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        if (direct)
            return launch<false, true>(ma, mb, p, ctx->sm_count, s);
        return launch_ff(ma, mb, p, ctx->sm_count, s);
    };
    st = body();
    if (x_shared && st == AXB_OK)
        *d.x_prepared = xq; // the caller's to free
    else if (!x_ready)
        axb_free(xq, s); // stream-ordered: reusable after the GEMM
    axb_free(wq, s);
    if (st == AXB_OK) {
        *done = true;
        last_path() = 4;
    }
    return st;
}

int &last_path() {
    static thread_local int v = 0;
    return v;
}

void invalidate_const_cache(const void *ptr, size_t bytes) {
    if (g_split_n.load(std::memory_order_acquire) == 0 || !ptr)
        return;
    const char *lo = static_cast<const char *>(ptr), *hi = lo + bytes;
    std::vector<void *> dead;
    {
        std::lock_guard<std::mutex> lk(g_split_mu);
        for (size_t i = 0; i < g_split.size();) {
            if (g_split[i].lo < hi && lo < g_split[i].hi) {
                dead.push_back(g_split[i].pieces);
                g_split_stats.bytes -= g_split[i].bytes;
                g_split_stats.invalidations++;
                g_split.erase(g_split.begin() + i);
            } else {
                ++i;
            }
        }
        g_split_stats.entries = g_split.size();
        g_split_n.store((int)g_split.size(), std::memory_order_release);
    }
    for (void *q : dead)
        axb_free(q, nullptr); // stream-ordered: GEMMs already queued on the backend stream still read valid pieces
}
bool const_cache_active() { return g_split_n.load(std::memory_order_acquire) != 0; }
axb_const_cache_stats const_cache_stats() {
    std::lock_guard<std::mutex> lk(g_split_mu);
    return g_split_stats;
}
PresplitA &presplit_a_hint() {
    static thread_local PresplitA h;
    return h;
}

axb_status try_tcgen05(const Desc &d, int adt, int bdt, int cdt, DeviceCtx *ctx, cudaStream_t s, bool *done) {
    *done = false;
    if (adt == AXB_F32 && bdt == AXB_F32 && cdt == AXB_F32)
        return try_split_f32(d, ctx, s, done);
    if (!((adt == AXB_BF16 || adt == AXB_F16) && adt == bdt && cdt == adt))
        return AXB_OK;
    if (d.M * d.N * d.K < 64 * 64 * 64 || d.M >= (1ll << 31) || d.N >= (1ll << 31) || d.K >= (1ll << 31) || d.batch >= (1ll << 31))
        return AXB_OK; // tiny problems stay on the SIMT kernel
    Operand oa, ob;
    if (!classify(d.a_rs, d.a_cs, d.M, d.K, &oa) || !classify(d.b_cs, d.b_rs, d.N, d.K, &ob))
        return AXB_OK;
    int64_t a_bs, b_bs, c_bs, bias_bs = 0;
    int a_mult, b_mult, c_mult, bias_mult = 0;
    if (!collapse_batch(d, d.a_bs, &a_bs, &a_mult) || !collapse_batch(d, d.b_bs, &b_bs, &b_mult) ||
        !collapse_batch(d, d.c_bs, &c_bs, &c_mult))
        return AXB_OK;
    if (d.bias && !collapse_batch(d, d.bias_bs, &bias_bs, &bias_mult))
        return AXB_OK;
    if (d.batch > 1 && !c_mult)
        return AXB_OK;
    auto aligned16 = [](int64_t v) { return v % 16 == 0; };
    if (!aligned16((int64_t)reinterpret_cast<uintptr_t>(d.a)) || !aligned16((int64_t)reinterpret_cast<uintptr_t>(d.b)) ||
        !aligned16(oa.ld) || !aligned16(ob.ld) || !aligned16(a_bs) || !aligned16(b_bs))
        return AXB_OK;
    if (d.c_rs < 0 || d.c_cs < 0)
        return AXB_OK;
    const drv::Api *api = drv::api();
    if (!api)
        return AXB_ERR_CUDA;
    // degenerate extents: a dim of extent 1 may carry any stride; give TMA a legal one
    const uint64_t a_ld = (uint64_t)std::max<int64_t>(oa.ld, 16), b_ld = (uint64_t)std::max<int64_t>(ob.ld, 16);
    const uint64_t a_b = (uint64_t)std::max<int64_t>(a_bs, 16), b_b = (uint64_t)std::max<int64_t>(b_bs, 16);
    const uint64_t abatch = a_mult ? (uint64_t)d.batch : 1, bbatch = b_mult ? (uint64_t)d.batch : 1;
    const int bn = pick_bn(d.N);
    const bool pair = use_pair(d, bn, ctx->sm_count);
    CUtensorMap ma, mb;
    if (!oa.mn_major)
        AXB_TRY(make_map(&ma, api, adt, d.a, (uint64_t)d.K, (uint64_t)d.M, abatch, a_ld, a_b, BK, BM));
    else
        AXB_TRY(make_map(&ma, api, adt, d.a, (uint64_t)d.M, (uint64_t)d.K, abatch, a_ld, a_b, 64, BK));
    if (!ob.mn_major)
        AXB_TRY(make_map(&mb, api, bdt, d.b, (uint64_t)d.K, (uint64_t)d.N, bbatch, b_ld, b_b, BK, pair ? 128u : (uint32_t)bn));
    else
        AXB_TRY(make_map(&mb, api, bdt, d.b, (uint64_t)d.N, (uint64_t)d.K, bbatch, b_ld, b_b, 64, BK));
    TcParams p;
    memset(&p, 0, sizeof p);
    p.M = d.M; p.N = d.N; p.K = d.K; p.batch = d.batch;
    p.tiles_m = pair ? (int)((d.M + 255) / 256) : (int)((d.M + BM - 1) / BM);
    p.bn = bn;
    p.tiles_n = (int)((d.N + bn - 1) / bn);
    p.num_tiles = (int64_t)p.tiles_m * p.tiles_n * d.batch;
    p.c = d.c; p.c_rs = d.c_rs; p.c_cs = d.c_cs; p.c_bs = c_bs;
    p.bias = d.bias; p.bias_rs = d.bias_rs; p.bias_cs = d.bias_cs; p.bias_bs = bias_mult ? bias_bs : 0;
    p.bias_dtype = d.bias_dtype; p.act = d.act; p.out_dtype = cdt;
    p.a_bmult = a_mult; p.b_bmult = b_mult;
    static const int group_m_env = getenv("AXB_GEMM_GROUP_M") ? std::max(1, atoi(getenv("AXB_GEMM_GROUP_M"))) : GROUP_M; // tuning knob
    p.group_m = group_m_env;
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format F32 = 1 @4, a/b format (F16 = 0,
    // BF16 = 1) @7/@10, a/b major (MN = 1) @15/@16, N>>3 @17, M>>4 @24
    const uint32_t fmt = adt == AXB_BF16 ? 1u : 0u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((oa.mn_major ? 1u : 0u) << 15) | ((ob.mn_major ? 1u : 0u) << 16) |
              ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((pair ? 256 : BM) >> 4) << 24);
    axb_status st;
    if (pair) {
        if (!oa.mn_major && !ob.mn_major)
            st = launch_pair<false, false>(ma, mb, p, ctx->sm_count, s);
        else if (!oa.mn_major && ob.mn_major)
            st = launch_pair<false, true>(ma, mb, p, ctx->sm_count, s);
        else if (oa.mn_major && !ob.mn_major)
            st = launch_pair<true, false>(ma, mb, p, ctx->sm_count, s);
        else
            st = launch_pair<true, true>(ma, mb, p, ctx->sm_count, s);
    } else if (!oa.mn_major && !ob.mn_major)
        st = launch<false, false>(ma, mb, p, ctx->sm_count, s);
    else if (!oa.mn_major && ob.mn_major)
        st = launch<false, true>(ma, mb, p, ctx->sm_count, s);
    else if (oa.mn_major && !ob.mn_major)
        st = launch<true, false>(ma, mb, p, ctx->sm_count, s);
    else
        st = launch<true, true>(ma, mb, p, ctx->sm_count, s);
    if (st == AXB_OK) {
        *done = true;
        last_path() = pair ? 2 : 1;
    }
    return st;
}

} // namespace gemm
} // namespace axb
