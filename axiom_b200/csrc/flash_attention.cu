// flash_attention.cu — scaled dot-product attention as ONE tcgen05 kernel: S = Q K^T and O = P V
// on the 5th-generation tensor cores, online softmax between them in registers, the score matrix never leaves the SM.
//
// Replaces cpu_flash_attention (src/backends/cpu/flash_attention_cpu.cpp:18-131: 64x64 tiles, sgemm + std::exp, the
// online-softmax recurrence m / l / alpha) and the Metal kernel (src/backends/metal/flash_attention.metal:231-275) behind
// the ops::scaled_dot_product_attention hook (src/tensor/operations.cpp:4031). Semantics kept: scores scaled by `scale`,
// causal masks kv > q, explicit mask byte != 0 drops the position (flash_attention_cpu.cpp:77-105), a fully masked row is
// NaN (exp(-inf - -inf)), O = acc / l.
//
// Data path. A one-launch prologue (fa_prep_kernel) turns the strided {B,H,L,D} operands into K-major bf16 tiles
// [B*H][L][NP*64]: fp32 inputs as the exact three-piece split x = hi + mid + lo (NP = 3, gemm_tcgen05.cu), bf16 / fp16
// inputs as they are (NP = 1), head_dim zero-padded to 64. The main kernel runs one CTA per (batch*head, 256 query rows):
//   warp 8        TMA producer : Q tiles once, then K and V blocks of 64 keys through two 2-stage rings (128-byte swizzle)
//   warp 9        MMA issuer   : S_t = Q_t K^T (A, B from shared memory) and O_t += P_t V (A = P straight from TENSOR
//                                MEMORY, B = V MN-major from shared memory); fp32 accuracy = six bf16 piece products per
//                                MMA step, smallest first, in one fp32 accumulator
//   warps 0..7    softmax      : two query tiles of 128 rows (warps 0-3 / 4-7), thread = row: tcgen05.ld of the 64 scores,
//                                running max with LAZY rescaling (the accumulator is only rescaled when the max grows by
//                                more than 2^8, so the TMEM round trip of O is rare), exp2, row sum, split of P into bf16
//                                pieces written back to TMEM with tcgen05.st. P is carried as TWO pieces (hi + mid, 16
//                                significand bits): its rounding errors are independent per key and average out over
//                                the row (2^-17 / sqrt(keys) relative in O), so the P_lo . V_hi product is dropped
// TMEM per query tile (256 of the 512 columns): TWO score buffers of 64 columns and O (64); the pieces of P(j) are
// written over the score buffer they were computed from, once its 64 scores sit in registers. The MMA issuer runs two key blocks ahead of the softmax: Q K^T of block j+2 is issued right after P V of block
// j (it reuses that score buffer; tcgen05.mma executes in issue order), so the softmax warps always find the next scores
// ready and the tensor pipe never waits for one exp round trip (the first version was serialised per tile: softmax -> P V
// -> Q K^T -> softmax, ncu: tensor pipe 39 % active, 0.229 ms on the whisper shape).
// head_dim <= 64 (whisper: 64); wider heads keep the three-kernel path in attention.cu.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "axb_internal.h"
#include "driver.h"
#include "reduce_common.cuh"

namespace axb {
namespace fa {

namespace {

constexpr int BQ = 128, BKV = 64, DP = 64;
constexpr int NQT = 2;            // query tiles per CTA
constexpr int KST = 2, VST = 2;   // ring depths
constexpr int NUM_THREADS = 320;  // 8 softmax warps, TMA producer, MMA issuer
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_S = 0 /* + 64 * buf */, COL_O = 192, COL_TILE = 256; // columns 128..191 of a tile are free
constexpr int Q_PIECE = BQ * DP * 2;  // 16 KiB
constexpr int KV_PIECE = BKV * DP * 2; // 8 KiB
constexpr float LAZY = 8.0f;      // log2 units: P stays below 2^8

struct Params {
    char *out;
    int64_t o_sb, o_sh, o_sq, o_sd; // bytes
    int32_t out_dtype;
    const uint8_t *mask;
    int64_t m_sb, m_sh, m_sq, m_sk;
    int32_t H, Lq, Lk, Dv;
    float scale_log2; // scale * log2(e)
    int32_t causal;
    uint32_t idesc_qk, idesc_pv;
};

// ---- PTX wrappers (same conventions as gemm_tcgen05.cu) ---------------------------------------------------------------
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]   (cute SM100_MMA_F16BF16_TS: A is K-major, 128 lanes x 8 columns of packed pairs per step)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
                   "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
                   "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// UMMA shared-memory descriptor, SWIZZLE_128B (see gemm_tcgen05.cu umma_desc)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// one lane of a fully converged warp (the warp runs the role's control flow uniformly, so descriptors and barrier addresses
// stay in uniform registers; issuing from inside `if (lane == 0)` cost ~10 extra instructions per MMA: R2UR moves and an
// ELECT retry loop per tcgen05.mma — more than the 32-48 cycles a 128x64x16 MMA takes to execute)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// piece (0 = hi, 1 = mid, 2 = lo) of A and B for combination c, smallest products first (gemm_tcgen05.cu)
__device__ __forceinline__ int piece_a(int c) { return (0x001120 >> (4 * c)) & 0xF; }
__device__ __forceinline__ int piece_b(int c) { return (0x010102 >> (4 * c)) & 0xF; }

// ---- prologue: strided {B,H,L,D} -> [B*H][L][NP*64] half tiles --------------------------------------------------------
struct PrepOne {
    const char *src;
    uint16_t *dst;
    int64_t sb, sh, sl, sd; // bytes
    int32_t L, D;
};
struct PrepParams {
    PrepOne t[3];
    int32_t H, in_dtype, np;
};
// grid (ceil(Lmax / 32), B*H, 3), block 256: a thread converts 8 consecutive d of one row
__global__ void __launch_bounds__(256) fa_prep_kernel(const __grid_constant__ PrepParams p) {
    const PrepOne &t = p.t[blockIdx.z];
    const int r = blockIdx.x * 32 + (threadIdx.x >> 3), d0 = (threadIdx.x & 7) * 8;
    if (r >= t.L)
        return;
    const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
    const char *row = t.src + b * t.sb + h * t.sh + (int64_t)r * t.sl;
    uint16_t *drow = t.dst + ((int64_t)bh * t.L + r) * (p.np * DP) + d0;
    if (p.in_dtype == AXB_F32) {
        float v[8];
        if (t.sd == 4 && d0 + 8 <= t.D && ((reinterpret_cast<uintptr_t>(row) + d0 * 4) & 15) == 0) {
            const float4 a = *reinterpret_cast<const float4 *>(row + d0 * 4), c = *reinterpret_cast<const float4 *>(row + d0 * 4 + 16);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = d0 + i < t.D ? *reinterpret_cast<const float *>(row + (int64_t)(d0 + i) * t.sd) : 0.0f;
        }
        uint32_t hh[4], mm[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            red::split3_pair(v[2 * i], v[2 * i + 1], hh[i], mm[i], ll[i]);
        *reinterpret_cast<uint4 *>(drow) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4 *>(drow + DP) = make_uint4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<uint4 *>(drow + 2 * DP) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    } else {
        uint16_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            v[i] = d0 + i < t.D ? *reinterpret_cast<const uint16_t *>(row + (int64_t)(d0 + i) * t.sd) : (uint16_t)0;
        *reinterpret_cast<uint4 *>(drow) = make_uint4(v[0] | ((uint32_t)v[1] << 16), v[2] | ((uint32_t)v[3] << 16),
                                                      v[4] | ((uint32_t)v[5] << 16), v[6] | ((uint32_t)v[7] << 16));
    }
}

// ---- main kernel --------------------------------------------------------------------------------------------------------
// NP: pieces per operand (3 = fp32 inputs, 1 = bf16 / fp16 inputs). F16: fp16 operands (P is packed as fp16).
template <int NP, bool F16>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
          const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ Params p) {
    constexpr int NC = NP == 3 ? 6 : 1; // piece combinations per MMA step
    constexpr int Q_TILE = NP * Q_PIECE, KV_STAGE = NP * KV_PIECE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem = (raw + 1023u) & ~1023u;
    const uint32_t q_base = smem, k_base = q_base + NQT * Q_TILE, v_base = k_base + KST * KV_STAGE;
    const uint32_t bar_base = v_base + VST * KV_STAGE;
    auto q_full = [&](int t) { return bar_base + 8u * t; };
    auto k_full = [&](int s) { return bar_base + 8u * (2 + s); };
    auto k_empty = [&](int s) { return bar_base + 8u * (4 + s); };
    auto v_full = [&](int s) { return bar_base + 8u * (6 + s); };
    auto v_empty = [&](int s) { return bar_base + 8u * (8 + s); };
    auto s_full = [&](int t, int buf) { return bar_base + 8u * (10 + 2 * t + buf); }; // one per score buffer: lead <= 1 phase
    auto p_full = [&](int t) { return bar_base + 8u * (14 + t); };
    auto o_done = [&](int t) { return bar_base + 8u * (16 + t); };
    const uint32_t tmem_slot = bar_base + 8u * 18;
    uint32_t *tmem_slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    // causal: the CTAs with the most key blocks (largest q0) are scheduled first
    const int q0 = (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * (NQT * BQ);
    // number of key blocks this CTA visits: all of them, or (causal) up to the block holding its last query row
    int nkv = (p.Lk + BKV - 1) / BKV;
    if (p.causal) {
        const int q_last = min(p.Lq, q0 + NQT * BQ) - 1;
        nkv = min(nkv, q_last / BKV + 1);
    }
    // a tile whose rows all lie beyond Lq does no work at all (Lq <= 128: tile 1)
    const bool tile_on[NQT] = {true, q0 + BQ < p.Lq};

    if (threadIdx.x == 0) {
        for (int t = 0; t < NQT; ++t) {
            mbar_init(q_full(t), 1);
            mbar_init(s_full(t, 0), 1);
            mbar_init(s_full(t, 1), 1);
            mbar_init(p_full(t), 4); // one arrive per softmax warp of the tile
            mbar_init(o_done(t), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(v_empty(s), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 8) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int t = 0; t < NQT; ++t) {
                if (!tile_on[t])
                    continue;
                mbar_expect_tx(q_full(t), Q_TILE);
#pragma unroll
                for (int pc = 0; pc < NP; ++pc)
                    tma_load_3d(q_base + t * Q_TILE + pc * Q_PIECE, &tma_q, q_full(t), pc * DP, q0 + t * BQ, bh);
            }
            for (int j = 0; j < nkv; ++j) {
                const int ks = j % KST, vs = j % VST;
                mbar_wait(k_empty(ks), ((uint32_t)(j / KST) & 1u) ^ 1u);
                mbar_expect_tx(k_full(ks), KV_STAGE);
#pragma unroll
                for (int pc = 0; pc < NP; ++pc)
                    tma_load_3d(k_base + ks * KV_STAGE + pc * KV_PIECE, &tma_k, k_full(ks), pc * DP, j * BKV, bh);
                mbar_wait(v_empty(vs), ((uint32_t)(j / VST) & 1u) ^ 1u);
                mbar_expect_tx(v_full(vs), KV_STAGE);
#pragma unroll
                for (int pc = 0; pc < NP; ++pc)
                    tma_load_3d(v_base + vs * KV_STAGE + pc * KV_PIECE, &tma_v, v_full(vs), pc * DP, j * BKV, bh);
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer: the whole warp walks the schedule, one elected lane issues =====
        {
            // descriptors: base + (byte offset >> 4) in the 14-bit start-address field (shared memory < 256 KiB: no carry)
            const uint64_t qdesc0 = umma_desc(q_base, 16, 1024), kdesc0 = umma_desc(k_base, 16, 1024), vdesc0 = umma_desc(v_base, 8192, 1024);
            auto qk = [&](int t, int ks, int buf, uint32_t bar_s, uint32_t bar_k /* 0: none */) { // S_t[buf] = Q_t K^T
                const uint32_t d = tmem_base + t * COL_TILE + COL_S + buf * 64;
                const uint64_t qa = qdesc0 + (uint64_t)((t * Q_TILE) >> 4), kb = kdesc0 + (uint64_t)((ks * KV_STAGE) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint64_t pa = qa + (uint64_t)(((NP == 3 ? piece_a(c) : 0) * Q_PIECE) >> 4);
                        const uint64_t pb = kb + (uint64_t)(((NP == 3 ? piece_b(c) : 0) * KV_PIECE) >> 4);
#pragma unroll
                        for (int k = 0; k < DP / 16; ++k)
                            tc_mma_ss(d, pa + (uint64_t)(k * 2), pb + (uint64_t)(k * 2), p.idesc_qk, (c | k) != 0);
                    }
                    tc_commit(bar_s);
                    if (bar_k)
                        tc_commit(bar_k);
                }
                __syncwarp();
            };
            auto pv = [&](int t, int vs, int j, uint32_t bar_o, uint32_t bar_v /* 0: none */) { // O_t += P_t(j) V
                const int buf = j & 1;
                const uint32_t tb = tmem_base + t * COL_TILE, d = tb + COL_O;
                // pieces of P(j) (hi, mid) sit over score buffer `buf`; five products, smallest first:
                // P_hi V_lo, P_mid V_mid, P_mid V_hi, P_hi V_mid, P_hi V_hi   (P_lo V_hi is below the P rounding noise)
                constexpr int NCP = NP == 3 ? 5 : 1;
                const uint32_t pp0 = tb + COL_S + buf * 64;
                const uint64_t vb = vdesc0 + (uint64_t)((vs * KV_STAGE) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < NCP; ++c) {
                        const int ap = NP == 3 ? ((0x00110 >> (4 * c)) & 0xF) : 0; // c = 0..4: 0,1,1,0,0
                        const int bp = NP == 3 ? ((0x01012 >> (4 * c)) & 0xF) : 0; // c = 0..4: 2,1,0,1,0
                        const uint32_t pa = pp0 + ap * 32;
                        const uint64_t pb = vb + (uint64_t)((bp * KV_PIECE) >> 4);
#pragma unroll
                        for (int k = 0; k < BKV / 16; ++k)
                            tc_mma_ts(d, pa + k * 8, pb + (uint64_t)(k * 128), p.idesc_pv, (j | c | k) != 0);
                    }
                    tc_commit(bar_o);
                    if (bar_v)
                        tc_commit(bar_v);
                }
                __syncwarp();
            };
            const int last_tile = tile_on[1] ? 1 : 0;
            // scores of the first two key blocks, both tiles
            for (int jj = 0; jj < 2 && jj < nkv; ++jj) {
                mbar_wait(k_full(jj % KST), 0);
                for (int t = 0; t <= last_tile; ++t) {
                    if (jj == 0)
                        mbar_wait(q_full(t), 0);
                    tc_fence_after();
                    qk(t, jj % KST, jj & 1, s_full(t, jj & 1), t == last_tile ? k_empty(jj % KST) : 0u);
                }
            }
            for (int j = 0; j < nkv; ++j) {
                const int vs = j % VST, ks2 = (j + 2) % KST;
                const bool more = j + 2 < nkv;
                for (int t = 0; t <= last_tile; ++t) {
                    mbar_wait(p_full(t), (uint32_t)j & 1u); // P_t(j) written (over the scores of block j)
                    if (t == 0)
                        mbar_wait(v_full(vs), (uint32_t)(j / VST) & 1u);
                    tc_fence_after();
                    pv(t, vs, j, o_done(t), t == last_tile ? v_empty(vs) : 0u);
                    if (more) { // scores of block j + 2 into the buffer P(j) is being read from: executes after P V above
                        if (t == 0)
                            mbar_wait(k_full(ks2), (uint32_t)((j + 2) / KST) & 1u);
                        tc_fence_after();
                        qk(t, ks2, j & 1, s_full(t, j & 1), t == last_tile ? k_empty(ks2) : 0u);
                    }
                }
            }
        }
    } else {
        // ===== softmax warps: tile = warp / 4, TMEM lane quarter = warp % 4, thread = query row =====
        const int t = warp >> 2, qd = warp & 3;
        if (tile_on[t]) {
            const int row = qd * 32 + lane;
            const int q = q0 + t * BQ + row;
            const bool row_ok = q < p.Lq;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + t * COL_TILE;
            const int b = bh / p.H, h = bh - b * p.H;
            const uint8_t *mrow = (p.mask && row_ok) ? p.mask + b * p.m_sb + h * p.m_sh + (int64_t)q * p.m_sq : nullptr;
            float m_used = -INFINITY, l = 0.0f;
            for (int j = 0; j < nkv; ++j) {
                const int buf = j & 1;
                mbar_wait(s_full(t, buf), (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                uint32_t s0[32], s1[32];
                tc_ld32(lane_addr + COL_S + buf * 64, s0);
                tc_ld32(lane_addr + COL_S + buf * 64 + 32, s1);
                tc_wait_ld();
                auto sr = [&](int i) { return __uint_as_float(i < 32 ? s0[i & 31] : s1[i & 31]); };
                const int kv0 = j * BKV;
                float x[64];
                float mx = -INFINITY;
                const bool edge = kv0 + BKV > p.Lk || (p.causal && kv0 + BKV - 1 > q) || mrow != nullptr;
                // x holds the RAW scores; the scale (> 0) is folded into the exponent: exp2(x * c - m), one FFMA per element
                if (!edge) {
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        x[i] = sr(i);
                        mx = fmaxf(mx, x[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const int kv = kv0 + i;
                        bool drop = kv >= p.Lk || (p.causal && kv > q);
                        if (mrow && kv < p.Lk)
                            drop = drop || mrow[(int64_t)kv * p.m_sk] != 0;
                        x[i] = drop ? -INFINITY : sr(i);
                        mx = fmaxf(mx, x[i]);
                    }
                }
                mx *= p.scale_log2;
                // lazy running max: keep the old reference point while the new maximum stays within 2^LAZY of it
                const bool grow = mx > m_used + LAZY || (m_used == -INFINITY && mx > -INFINITY);
                float alpha = 1.0f;
                if (grow) {
                    alpha = ex2(m_used - mx); // 0 on the first block (m_used = -inf)
                    m_used = mx;
                    l *= alpha;
                }
                // exponentials, row sum, pieces of P -> tensor memory (column c of a piece holds kv = 2c, 2c + 1)
                float sum = 0.0f;
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    uint32_t ph[16], pm[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e0 = ex2(fmaf(x[hb * 32 + 2 * i], p.scale_log2, -m_used)),
                                    e1 = ex2(fmaf(x[hb * 32 + 2 * i + 1], p.scale_log2, -m_used));
                        sum += e0 + e1;
                        if (NP == 3) { // hi = bf16(e), mid = bf16(e - hi): the first two steps of the exact split
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ph[i]) : "f"(e1), "f"(e0));
                            const float r0 = e0 - __uint_as_float(ph[i] << 16), r1 = e1 - __uint_as_float(ph[i] & 0xFFFF0000u);
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pm[i]) : "f"(r1), "f"(r0));
                        } else if (F16) {
                            ph[i] = (uint32_t)red::f2h(e0) | ((uint32_t)red::f2h(e1) << 16);
                        } else {
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ph[i]) : "f"(e1), "f"(e0));
                        }
                    }
                    // 16 columns (32 keys) per half and piece
                    // store 16 columns with a x16 store
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                 ::"r"(lane_addr + COL_S + buf * 64 + hb * 16), "r"(ph[0]), "r"(ph[1]), "r"(ph[2]), "r"(ph[3]), "r"(ph[4]), "r"(ph[5]),
                                   "r"(ph[6]), "r"(ph[7]), "r"(ph[8]), "r"(ph[9]), "r"(ph[10]), "r"(ph[11]), "r"(ph[12]), "r"(ph[13]),
                                   "r"(ph[14]), "r"(ph[15])
                                 : "memory");
                    if (NP == 3) {
                        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                                     ::"r"(lane_addr + COL_S + buf * 64 + 32 + hb * 16), "r"(pm[0]), "r"(pm[1]), "r"(pm[2]), "r"(pm[3]), "r"(pm[4]),
                                       "r"(pm[5]), "r"(pm[6]), "r"(pm[7]), "r"(pm[8]), "r"(pm[9]), "r"(pm[10]), "r"(pm[11]), "r"(pm[12]),
                                       "r"(pm[13]), "r"(pm[14]), "r"(pm[15])
                                     : "memory");
                    }
                }
                l += sum;
                // O_t: every phase of o_done is consumed in order (a parity wait is only meaningful one phase ahead), but as LATE
                // as possible — P(j) went over score buffer j, which P V(j-1) does not touch, so only the rescale of O needs
                // P V(j-1) to be complete, and by now it has had this whole score block of time (waiting here BEFORE the
                // exponentials cost 45 % of the softmax warps' samples). Rare: the reference point moved -> TMEM round trip.
                if (j > 0) {
                    mbar_wait(o_done(t), (uint32_t)(j - 1) & 1u);
                    if (__any_sync(0xFFFFFFFFu, grow)) { // warp-uniform: tcgen05.ld / st are warp-collective
                        tc_fence_after();
#pragma unroll
                        for (int hb = 0; hb < 2; ++hb) {
                            uint32_t o[32];
                            tc_ld32(lane_addr + COL_O + hb * 32, o);
                            tc_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tc_st32(lane_addr + COL_O + hb * 32, o);
                        }
                    }
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(p_full(t));
            }
            // epilogue: O / l -> global memory
            mbar_wait(o_done(t), (uint32_t)(nkv - 1) & 1u);
            tc_fence_after();
            const float inv = l > 0.0f ? __fdiv_rn(1.0f, l) : 0.0f; // flash_attention_cpu.cpp:138 (a NaN l gives 0 * NaN = NaN)
            char *orow = p.out + b * p.o_sb + h * p.o_sh + (int64_t)q * p.o_sq;
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                uint32_t o[32];
                tc_ld32(lane_addr + COL_O + hb * 32, o);
                tc_wait_ld();
                if (!row_ok)
                    continue;
                if (p.out_dtype == AXB_F32 && p.o_sd == 4 && hb * 32 + 32 <= p.Dv && (reinterpret_cast<uintptr_t>(orow) & 15) == 0) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4 *>(orow + (hb * 32 + i) * 4) =
                            make_float4(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv, __uint_as_float(o[i + 2]) * inv,
                                        __uint_as_float(o[i + 3]) * inv);
                    continue;
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int d = hb * 32 + i;
                    if (d >= p.Dv)
                        break;
                    const float y = __uint_as_float(o[i]) * inv;
                    char *dst = orow + (int64_t)d * p.o_sd;
                    if (p.out_dtype == AXB_F32)
                        *reinterpret_cast<float *>(dst) = y;
                    else if (p.out_dtype == AXB_BF16)
                        *reinterpret_cast<uint16_t *>(dst) = red::f2bf(y);
                    else
                        *reinterpret_cast<uint16_t *>(dst) = red::f2h(y);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

axb_status make_map(CUtensorMap *map, const drv::Api *api, bool f16, const void *base, uint64_t d0, uint64_t d1, uint64_t d2,
                    uint32_t b0, uint32_t b1) {
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstr[2] = {d0 * 2, d0 * d1 * 2};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = api->TensorMapEncodeTiled(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                                           const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(AXB_ERR_CUDA, "cuTensorMapEncodeTiled failed: %s", drv::err_str(r));
    return AXB_OK;
}

template <int NP, bool F16>
axb_status launch(const CUtensorMap &mq, const CUtensorMap &mk, const CUtensorMap &mv, const Params &p, int64_t BH, cudaStream_t s) {
    constexpr int SMEM = NQT * NP * Q_PIECE + (KST + VST) * NP * KV_PIECE + 1024 + 256;
    static bool configured = false;
    if (!configured) {
        AXB_CUDA(cudaFuncSetAttribute(fa_kernel<NP, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        configured = true;
    }
    const dim3 grid((unsigned)((p.Lq + NQT * BQ - 1) / (NQT * BQ)), (unsigned)BH);
    fa_kernel<NP, F16><<<grid, NUM_THREADS, SMEM, s>>>(mq, mk, mv, p);
    AXB_CHECK_LAUNCH();
    return AXB_OK;
}

} // namespace

bool supported(int32_t dtype, int64_t D, int64_t Dv, int64_t BH, int64_t Lq, int64_t Lk) {
    const bool off = getenv("AXB_SDPA_NO_FLASH") != nullptr; // A/B switch (read per call): the three-kernel path of attention.cu
    return !off && (dtype == AXB_F32 || dtype == AXB_BF16 || dtype == AXB_F16) && D >= 1 && D <= DP && Dv >= 1 && Dv <= DP &&
           BH <= 65535 && Lq < (1ll << 30) && Lk < (1ll << 30);
}

// q {B,H,Lq,D}, k {B,H,Lk,D}, v {B,H,Lk,Dv}, out {B,H,Lq,Dv}; all of one float dtype; shapes already validated.
axb_status run(const axb_tensor *q, const axb_tensor *k, const axb_tensor *v, const uint8_t *mask, const int64_t mask_strides[4],
               float scale, int is_causal, const axb_tensor *out, DeviceCtx *ctx, cudaStream_t s) {
    (void)ctx;
    const int64_t B = q->shape[0], H = q->shape[1], Lq = q->shape[2], Lk = k->shape[2], Dv = v->shape[3];
    const int64_t BH = B * H;
    const int dt = q->dtype;
    const int np = dt == AXB_F32 ? 3 : 1;
    const drv::Api *api = drv::api();
    if (!api)
        return AXB_ERR_CUDA;
    const size_t qb = (size_t)BH * Lq * np * DP * 2, kb = (size_t)BH * Lk * np * DP * 2;
    void *ws = nullptr;
    AXB_TRY(axb_alloc(qb + 2 * kb, s, &ws));
    uint16_t *qp = static_cast<uint16_t *>(ws), *kp = qp + qb / 2, *vp = kp + kb / 2;
    PrepParams pp;
    memset(&pp, 0, sizeof pp);
    pp.H = (int)H; pp.in_dtype = dt; pp.np = np;
    const axb_tensor *src[3] = {q, k, v};
    uint16_t *dst[3] = {qp, kp, vp};
    for (int i = 0; i < 3; ++i) {
        pp.t[i].src = tensor_ptr(*src[i]);
        pp.t[i].dst = dst[i];
        pp.t[i].sb = src[i]->strides[0]; pp.t[i].sh = src[i]->strides[1]; pp.t[i].sl = src[i]->strides[2]; pp.t[i].sd = src[i]->strides[3];
        pp.t[i].L = (int)src[i]->shape[2]; pp.t[i].D = (int)src[i]->shape[3];
    }
    const int64_t lmax = std::max(Lq, Lk);
    fa_prep_kernel<<<dim3((unsigned)((lmax + 31) / 32), (unsigned)BH, 3), 256, 0, s>>>(pp);
    axb_status st = AXB_OK;
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            st = fail(AXB_ERR_CUDA, "axb_sdpa: prologue launch failed: %s", cudaGetErrorString(e));
        else
            count_launch();
    }
    CUtensorMap mq, mk, mv;
    const bool f16 = dt == AXB_F16;
    if (st == AXB_OK)
        st = make_map(&mq, api, f16, qp, (uint64_t)(np * DP), (uint64_t)Lq, (uint64_t)BH, DP, BQ);
    if (st == AXB_OK)
        st = make_map(&mk, api, f16, kp, (uint64_t)(np * DP), (uint64_t)Lk, (uint64_t)BH, DP, BKV);
    if (st == AXB_OK)
        st = make_map(&mv, api, f16, vp, (uint64_t)(np * DP), (uint64_t)Lk, (uint64_t)BH, DP, BKV);
    if (st == AXB_OK) {
        Params p;
        memset(&p, 0, sizeof p);
        p.out = tensor_ptr(*out);
        p.o_sb = out->strides[0]; p.o_sh = out->strides[1]; p.o_sq = out->strides[2]; p.o_sd = out->strides[3];
        p.out_dtype = out->dtype;
        p.mask = mask;
        if (mask) {
            p.m_sb = mask_strides[0]; p.m_sh = mask_strides[1]; p.m_sq = mask_strides[2]; p.m_sk = mask_strides[3];
        }
        p.H = (int)H; p.Lq = (int)Lq; p.Lk = (int)Lk; p.Dv = (int)Dv;
        p.scale_log2 = scale * 1.4426950408889634f;
        p.causal = is_causal ? 1 : 0;
        // instruction descriptor (cute/arch/mma_sm100_desc.hpp): c_format F32 = 1 @4, a/b format (F16 = 0, BF16 = 1) @7/@10,
        // b major (MN = 1) @16, N >> 3 @17, M >> 4 @24
        const uint32_t fmt = f16 ? 0u : 1u;
        p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BKV >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
        p.idesc_pv = p.idesc_qk | (1u << 16); // N = 64 (Dv padded) as well, B = V is MN-major
        if (np == 3)
            st = launch<3, false>(mq, mk, mv, p, BH, s);
        else if (f16)
            st = launch<1, true>(mq, mk, mv, p, BH, s);
        else
            st = launch<1, false>(mq, mk, mv, p, BH, s);
    }
    axb_free(ws, s);
    return st;
}

} // namespace fa
} // namespace axb
