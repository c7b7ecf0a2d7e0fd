// half_lut.cu — fused chains of ONE half-precision tensor as a table lookup.
//
// A program whose only tensor operand is one bf16 / fp16 tensor (every other operand an immediate) and whose result is a
// 2-byte type is a function of 16 bits: it has 65 536 possible results. For large tensors with a transcendental in the chain
// (gelu, exp, tanh, sigmoid, silu, erf, log, sin, cos, tan, pow, cbrt) the specialised chain kernel is instruction-bound —
// every element is widened to fp32, pushed through the reference's Highway polynomial (csrc/hwy_math.cuh, kept bit-identical
// to src/backends/cpu/simd/hwy_kernels.cc on all 65 536 inputs) and rounded once: 35 (exp) to 79 (gelu) instructions per
// element, 21-43 % of the HBM roofline (profiles/r1_ncu_summary_s7_elementwise.txt). Here the SAME kernel is run once over
// the 65 536 input patterns to build a 128 KiB table (kept per program, per device), and the tensor is then streamed through
// a persistent kernel that holds the table in shared memory: one 128-bit load, eight shared-memory lookups, one 128-bit
// store per eight elements. Results are bit-identical to the chain kernel by construction — the table IS its output.
//
// Falls back (returns *done = false) for anything else: several tensor operands, strided / broadcast operands, small
// tensors (the table costs 128 KiB of L2 -> shared traffic per SM), cheap chains that are already bandwidth-bound.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "axb_internal.h"

namespace axb {
namespace lut {

namespace {

constexpr int kThreads = 1024;
constexpr int kUnroll = 4;                  // independent 128-bit loads in flight per thread
constexpr size_t kTableBytes = 65536 * 2;   // one 2-byte result per input pattern
constexpr int64_t kMinElements = 1ll << 21; // below this the chain kernel wins (table fill + launch of a full grid)

__global__ void __launch_bounds__(256) iota16_kernel(uint16_t *p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 65536)
        p[i] = (uint16_t)i;
}

__global__ void __launch_bounds__(kThreads, 1) lut_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, const uint4 *__restrict__ table,
                                                          int64_t n_vec, const uint16_t *__restrict__ in_tail, uint16_t *__restrict__ out_tail,
                                                          int n_tail) {
    extern __shared__ uint4 tab4[];
    for (int i = threadIdx.x; i < (int)(kTableBytes / 16); i += kThreads)
        tab4[i] = table[i];
    __syncthreads();
    const uint16_t *tab = reinterpret_cast<const uint16_t *>(tab4);
    auto map2 = [&](uint32_t w) -> uint32_t { return (uint32_t)tab[w & 0xFFFFu] | ((uint32_t)tab[w >> 16] << 16); };
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n_vec; i += kUnroll * stride) {
        uint4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            v[u] = __ldcs(in + i + u * stride); // streamed once: evict-first keeps L2 for the table
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            uint4 r;
            r.x = map2(v[u].x);
            r.y = map2(v[u].y);
            r.z = map2(v[u].z);
            r.w = map2(v[u].w);
            __stcs(out + i + u * stride, r);
        }
    }
    for (; i < n_vec; i += stride) {
        const uint4 v = __ldcs(in + i);
        uint4 r;
        r.x = map2(v.x);
        r.y = map2(v.y);
        r.z = map2(v.z);
        r.w = map2(v.w);
        __stcs(out + i, r);
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail)
        out_tail[threadIdx.x] = tab[in_tail[threadIdx.x]];
}

bool is_half(int dt) { return dt == AXB_F16 || dt == AXB_BF16; }
bool is_table_result(int dt) { return dt == AXB_F16 || dt == AXB_BF16 || dt == AXB_I16 || dt == AXB_U16; }
bool costly(int op) {
    switch (op) {
    case AXB_OP_EXP: case AXB_OP_LOG: case AXB_OP_SIN: case AXB_OP_COS: case AXB_OP_TAN: case AXB_OP_ERF: case AXB_OP_TANH:
    case AXB_OP_SIGMOID: case AXB_OP_SILU: case AXB_OP_GELU: case AXB_OP_POW: case AXB_OP_CBRT: case AXB_OP_ATAN2: case AXB_OP_HYPOT:
        return true;
    default:
        return false;
    }
}

struct Entry {
    std::vector<unsigned char> key;
    void *table = nullptr;
};
std::mutex g_mu;
std::unordered_map<uint64_t, std::vector<Entry>> g_tables;
size_t g_count = 0;
axb_half_lut_stats g_stats{};
constexpr size_t kMaxTables = 64;
thread_local bool t_building = false;
bool g_smem_ready[64] = {};

void make_key(const axb_instr *prog, int n_instr, int in_dt, int device, std::vector<unsigned char> *key, uint64_t *hash) {
    key->clear();
    auto put = [&](const void *q, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(q);
        key->insert(key->end(), b, b + n);
    };
    const int32_t head[3] = {n_instr, in_dt, device};
    put(head, sizeof head);
    for (int i = 0; i < n_instr; ++i) {
        put(&prog[i].op, 4);
        put(&prog[i].dtype, 4);
        put(prog[i].src, 12);
        put(&prog[i].imm, 8);
    }
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : *key) {
        h ^= c;
        h *= 1099511628211ull;
    }
    *hash = h;
}

} // namespace

axb_status try_run(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out, cudaStream_t s,
                   DeviceCtx *ctx, bool *done) {
    *done = false;
    static const bool off = getenv("AXB_NO_HALF_LUT") != nullptr; // A/B switch
    if (off || t_building || n_inputs != 1 || !prog || n_instr < 1 || n_instr > AXB_MAX_INSTRS || !inputs || !out)
        return AXB_OK;
    const axb_tensor &x = inputs[0];
    if (!is_half(x.dtype) || !is_table_result(out->dtype) || prog[n_instr - 1].dtype != out->dtype || !x.base || !out->base)
        return AXB_OK;
    if (x.ndim != out->ndim || x.ndim < 0 || x.ndim > AXB_MAX_DIMS)
        return AXB_OK;
    for (int d = 0; d < x.ndim; ++d)
        if (x.shape[d] != out->shape[d])
            return AXB_OK;
    const int64_t n = numel(x);
    if (n < kMinElements || !is_contiguous(x) || !is_contiguous(*out))
        return AXB_OK;
    bool any_costly = false;
    for (int i = 0; i < n_instr; ++i) {
        any_costly = any_costly || costly(prog[i].op);
        for (int k = 0; k < 3; ++k) { // operands: input 0, an earlier value, an immediate, or none
            const int32_t sel = prog[i].src[k];
            if (sel >= 1)
                return AXB_OK;
        }
    }
    if (!any_costly)
        return AXB_OK;
    const char *ip = tensor_ptr(x);
    char *op = tensor_ptr(*out);
    if ((reinterpret_cast<uintptr_t>(ip) & 15u) || (reinterpret_cast<uintptr_t>(op) & 15u) || ctx->device < 0 || ctx->device >= 64)
        return AXB_OK;
    if (ctx->smem_optin < kTableBytes + 1024)
        return AXB_OK;

    // ---- table: built once per (program, input dtype, device) by the chain kernel itself ------------------------------------
    std::vector<unsigned char> key;
    uint64_t h = 0;
    make_key(prog, n_instr, x.dtype, ctx->device, &key, &h);
    void *table = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tables.find(h);
        if (it != g_tables.end())
            for (auto &e : it->second)
                if (e.key == key)
                    table = e.table;
        if (table)
            g_stats.hits++;
    }
    if (!table) {
        {
            std::lock_guard<std::mutex> lk(g_mu);
            if (g_count >= kMaxTables)
                return AXB_OK; // table memory is bounded; further programs keep the chain kernel
        }
        void *pat = nullptr;
        // cudaMalloc, not the pool: tables live for the life of the process and must not be handed out again
        AXB_CUDA(cudaMalloc(&table, kTableBytes));
        axb_status st = axb_alloc(kTableBytes, s, &pat);
        if (st == AXB_OK) {
            iota16_kernel<<<256, 256, 0, s>>>(static_cast<uint16_t *>(pat));
            if (cudaGetLastError() != cudaSuccess)
                st = fail(AXB_ERR_CUDA, "half LUT: pattern kernel launch failed");
            else
                count_launch();
        }
        if (st == AXB_OK) {
            const int64_t shape1[1] = {65536};
            const axb_tensor tin = make_contiguous_desc(pat, x.dtype, shape1, 1), tout = make_contiguous_desc(table, out->dtype, shape1, 1);
            t_building = true; // 65 536 elements are below kMinElements anyway; the flag documents the recursion
            st = axb_elementwise(prog, n_instr, &tin, 1, &tout, s);
            t_building = false;
        }
        if (pat)
            axb_free(pat, s);
        if (st != AXB_OK) {
            cudaFree(table);
            return st;
        }
        std::lock_guard<std::mutex> lk(g_mu);
        Entry e;
        e.key = key;
        e.table = table;
        g_tables[h].push_back(std::move(e));
        ++g_count;
        g_stats.tables = g_count;
        g_stats.builds++;
    }
    if (!g_smem_ready[ctx->device]) {
        AXB_CUDA(cudaFuncSetAttribute(lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTableBytes));
        g_smem_ready[ctx->device] = true;
    }
    const int64_t n_vec = n / 8;
    const int n_tail = (int)(n - n_vec * 8);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (n_vec + kThreads - 1) / kThreads));
    lut_kernel<<<blocks, kThreads, kTableBytes, s>>>(reinterpret_cast<const uint4 *>(ip), reinterpret_cast<uint4 *>(op),
                                                     static_cast<const uint4 *>(table), n_vec,
                                                     reinterpret_cast<const uint16_t *>(ip) + n_vec * 8,
                                                     reinterpret_cast<uint16_t *>(op) + n_vec * 8, n_tail);
    AXB_CHECK_LAUNCH();
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_stats.launches++;
    }
    *done = true;
    return AXB_OK;
}

} // namespace lut
} // namespace axb

extern "C" axb_status axb_get_half_lut_stats(axb_half_lut_stats *out) {
    AXB_REQUIRE(out, AXB_ERR_INVALID, "axb_get_half_lut_stats: null out");
    std::lock_guard<std::mutex> lk(axb::lut::g_mu);
    *out = axb::lut::g_stats;
    return AXB_OK;
}
