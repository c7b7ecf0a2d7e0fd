// reduce.cu — kernel family 2a: axis reductions (sum/mean/max/min/prod/any/all) and argmax/argmin.
//
// Replaces CPUReductionOperation (src/backends/cpu/cpu_operations.cpp:1307-1712) and CPUArgMax/ArgMin
// (:2157-2350). All of them are HBM-bound (C3 = 1 GiB in, 32-64 KiB out), so the design is about
// streaming the input once at full bandwidth:
//   * the strided input is split into KEPT dims (one output each) and REDUCED dims, each group collapsed;
//   * ROW kernel  — the reduced run is innermost/contiguous: one CTA (wide rows) or one warp (narrow rows) per
//     output, 128-bit streaming loads, 4 independent loads in flight per thread, warp-shuffle then shared-memory
//     two-level combine. Few outputs + long rows (full reductions) are split across CTAs into a workspace of
//     partials that a second tiny launch combines (no atomics: results are deterministic run to run);
//   * COL kernel  — the innermost KEPT dim is contiguous (e.g. sum over axis 0): threads map to consecutive
//     outputs (coalesced), a 32x8 CTA splits the reduced range over y and combines through shared memory;
//   * anything else takes the generic ROW kernel with per-element index arithmetic (correct, not fast).
// Numerics follow the reference: half inputs accumulate in fp32 and are rounded once (:1422-1424,1613-1614); mean =
// sum / N in the accumulation type (int: truncating); small ints wrap to their width; argmax/argmin keep the FIRST
// extreme, NaN (and values equal to lowest()/max()) never win, so an all-NaN row yields 0 (:2211-2228).
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "axb_internal.h"
#include "reduce_common.cuh"

namespace axb {
namespace red {

// ---------------------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------------------

// One output per group of TPR threads (TPR = 32: a warp, 8 rows per 256-thread CTA; TPR = 256: the CTA).
// blockIdx.y = split index (each split handles [y*chunk, (y+1)*chunk) of the reduced range).
template <typename In, typename Acc, class Op, int TPR, bool CONTIG>
__global__ void __launch_bounds__(256) reduce_rows_kernel(const __grid_constant__ Params p) {
    constexpr int ROWS = 256 / TPR;
    const int lane = threadIdx.x % TPR;
    const int64_t row = (int64_t)blockIdx.x * ROWS + threadIdx.x / TPR;
    const bool row_ok = row < p.n_out;
    Acc acc = Op::template identity<Acc>();
    if (row_ok) {
        const char *base = p.in + kept_offset(p, row);
        const int64_t r0 = (int64_t)blockIdx.y * p.chunk;
        const int64_t r1 = min(p.n_red, r0 + p.chunk);
        if (CONTIG) {
            constexpr int V = 16 / sizeof(In);
            const In *src = reinterpret_cast<const In *>(base);
            // head: scalar until 16-byte aligned
            int64_t r = r0;
            const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(src + r0) & 15) / sizeof(In));
            const int64_t head_end = min(r1, mis ? r0 + (V - mis) : r0);
            for (int64_t i = r + lane; i < head_end; i += TPR)
                acc = Op::combine(acc, Op::template load<In, Acc>(src[i]));
            r = head_end;
            const int64_t nvec = (r1 - r) / V;
            const uint4 *v = reinterpret_cast<const uint4 *>(src + r);
            int64_t i = lane;
            Acc a1 = Op::template identity<Acc>(), a2 = a1, a3 = a1;
            for (; i + 3 * TPR < nvec; i += 4 * TPR) { // 4 independent 128-bit loads in flight
                uint4 q0 = ldg_stream(v + i), q1 = ldg_stream(v + i + TPR), q2 = ldg_stream(v + i + 2 * TPR),
                      q3 = ldg_stream(v + i + 3 * TPR);
                acc = fold_vec<In, Acc, Op>(acc, q0);
                a1 = fold_vec<In, Acc, Op>(a1, q1);
                a2 = fold_vec<In, Acc, Op>(a2, q2);
                a3 = fold_vec<In, Acc, Op>(a3, q3);
            }
            for (; i < nvec; i += TPR)
                acc = fold_vec<In, Acc, Op>(acc, ldg_stream(v + i));
            acc = Op::combine(Op::combine(acc, a1), Op::combine(a2, a3));
            for (int64_t j = r + nvec * V + lane; j < r1; j += TPR)
                acc = Op::combine(acc, Op::template load<In, Acc>(src[j]));
        } else {
            for (int64_t r = r0 + lane; r < r1; r += TPR) {
                const In *e = reinterpret_cast<const In *>(base + reduced_offset(p, r));
                acc = Op::combine(acc, Op::template load<In, Acc>(*e));
            }
        }
    }
    acc = group_reduce<Acc, Op, TPR>(acc);
    if (row_ok && lane == 0) {
        if (p.split > 1 || p.peer_world > 1)
            reinterpret_cast<Acc *>(p.ws)[row * p.split + blockIdx.y] = acc;
        else
            finalize<Acc, Op>(p, row, acc);
    }
}

// Innermost kept dim contiguous: thread x owns VEC consecutive outputs, y strides over the reduced range.
template <typename In, typename Acc, class Op, int VEC>
__global__ void __launch_bounds__(256) reduce_cols_kernel(const __grid_constant__ Params p) {
    __shared__ Acc sm[8][32 * VEC + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t o0 = ((int64_t)blockIdx.x * 32 + tx) * VEC; // first output of this thread
    const int64_t r0 = (int64_t)blockIdx.y * p.chunk;
    const int64_t r1 = min(p.n_red, r0 + p.chunk);
    Acc acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k)
        acc[k] = Op::template identity<Acc>();
    if (o0 < p.n_out) {
        // VEC consecutive outputs share all outer kept coordinates (host guarantees inner kept extent % VEC == 0)
        const char *base = p.in + kept_offset(p, o0);
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const char *e = base + reduced_offset(p, r);
            if (VEC * sizeof(In) == 16) {
                uint4 q = ldg_stream(reinterpret_cast<const uint4 *>(e));
                const In *w = reinterpret_cast<const In *>(&q);
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc[k] = Op::combine(acc[k], Op::template load<In, Acc>(w[k]));
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc[k] = Op::combine(acc[k], Op::template load<In, Acc>(reinterpret_cast<const In *>(e)[k]));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k)
        sm[ty][tx * VEC + k] = acc[k];
    __syncthreads();
    if (ty == 0 && o0 < p.n_out) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            Acc a = sm[0][tx * VEC + k];
#pragma unroll
            for (int y = 1; y < 8; ++y)
                a = Op::combine(a, sm[y][tx * VEC + k]);
            if (p.split > 1 || p.peer_world > 1)
                reinterpret_cast<Acc *>(p.ws)[(o0 + k) * p.split + blockIdx.y] = a;
            else
                finalize<Acc, Op>(p, o0 + k, a);
        }
    }
}

// second pass of split reductions: one thread per output over `split` partials
// Sharded full reduction (the one exchange step of the path), fused into the finish kernel: thread 0 folds
// this GPU's split partials, then thread t < world STORES the shard's value into rank t's mailbox over NVLink and polls
// slot t of the own mailbox; thread 0 combines the `world` values in rank order (every rank computes the identical
// result) and finalises. One 64-bit word carries (call counter, fp32 value), so a reader never sees a torn pair; two
// banks selected by the counter's parity make the mailbox reusable without a reset: a rank can only reach call k + 2
// (which overwrites bank k & 1) after every rank has delivered call k + 1, i.e. after every rank has finished reading
// call k. No collective library call, no host round trip: the exchange costs one NVLink store + one polled load.
__device__ __forceinline__ void peer_store(uint64_t *addr, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t peer_load(const uint64_t *addr) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}
template <typename Acc, class Op> __device__ __forceinline__ void peer_finish(const Params &p) {
    if constexpr (sizeof(Acc) == 4 && std::is_floating_point<Acc>::value) {
        __shared__ float mine, vals[kMaxPeers];
        __shared__ int timed_out;
        if (threadIdx.x == 0) {
            const Acc *w = reinterpret_cast<const Acc *>(p.ws);
            Acc a = Op::template identity<Acc>();
            for (int s = 0; s < p.split; ++s)
                a = Op::combine(a, w[s]);
            mine = a;
            timed_out = 0;
        }
        __syncthreads();
        const int t = threadIdx.x;
        if (t < p.peer_world) {
            const uint32_t bank = (p.peer_seq & 1u) * kMaxPeers;
            peer_store(p.peer_box[t] + bank + p.peer_rank, ((uint64_t)p.peer_seq << 32) | __float_as_uint(mine));
            const uint64_t *slot = p.peer_box[p.peer_rank] + bank + t;
            uint64_t w = peer_load(slot);
            const long long t0 = clock64();
            while ((uint32_t)(w >> 32) != p.peer_seq) {
                // ~20 s (a late rank may still be compiling a kernel): a rank that never arrives must not hang the GPU.
                // The failure is REPORTED, not hidden in the value: the group's host-mapped error word receives the call
                // counter (the host turns it into AXB_ERR_CUDA on the next call / status query) and the result is NaN
                // for every op — a NaN operand alone would be dropped by max / min.
                if (clock64() - t0 > 40000000000ll) {
                    timed_out = 1;
                    if (p.peer_err)
                        atomicExch(p.peer_err, p.peer_seq);
                    break;
                }
                __nanosleep(64);
                w = peer_load(slot);
            }
            vals[t] = __uint_as_float((uint32_t)w);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            Acc tot = vals[0];
            for (int r = 1; r < p.peer_world; ++r)
                tot = Op::combine(tot, (Acc)vals[r]);
            if (timed_out)
                tot = __uint_as_float(0x7FC00000u);
            finalize<Acc, Op>(p, 0, tot);
        }
    }
}

template <typename Acc, class Op> __global__ void reduce_finish_kernel(const __grid_constant__ Params p) {
    if (p.peer_world > 1) {
        peer_finish<Acc, Op>(p);
        return;
    }
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.n_out)
        return;
    const Acc *w = reinterpret_cast<const Acc *>(p.ws) + o * p.split;
    Acc a = Op::template identity<Acc>();
    for (int s = 0; s < p.split; ++s)
        a = Op::combine(a, w[s]);
    finalize<Acc, Op>(p, o, a);
}

// ---- argmax / argmin ------------------------------------------------------------------------------
template <typename In, typename Cmp, bool IS_MAX, int TPR, bool CONTIG>
__global__ void __launch_bounds__(256) argreduce_rows_kernel(const __grid_constant__ Params p) {
    constexpr int ROWS = 256 / TPR;
    const int lane = threadIdx.x % TPR;
    const int64_t row = (int64_t)blockIdx.x * ROWS + threadIdx.x / TPR;
    const bool row_ok = row < p.n_out;
    typedef ArgOp<Cmp, IS_MAX> Op;
    ArgPair<Cmp> acc = Op::identity();
    if (row_ok) {
        const char *base = p.in + kept_offset(p, row);
        const int64_t r0 = (int64_t)blockIdx.y * p.chunk;
        const int64_t r1 = min(p.n_red, r0 + p.chunk);
        if (CONTIG) {
            // Each thread visits its elements in INCREASING index order, so "strictly better replaces" keeps the first
            // extremum exactly like the reference's serial loop (cpu_operations.cpp:2213-2222) at 3 instructions per
            // element (compare, select value, select 32-bit offset); NaN never compares better. Four independent scans
            // (one per in-flight 128-bit load) are merged with the tie -> lower index rule. Offsets are 32-bit relative
            // to a window start, windows of 2^30 elements keep them in range for any row length.
            constexpr int V = 16 / sizeof(In);
            const In *src = reinterpret_cast<const In *>(base);
            for (int64_t w0 = r0; w0 < r1; w0 += (int64_t)1 << 30) {
                const int64_t w1 = min(r1, w0 + ((int64_t)1 << 30));
                ArgScan<Cmp, IS_MAX> s0, s1, s2, s3;
                s0.init(); s1.init(); s2.init(); s3.init();
                const int64_t mis = (int64_t)((reinterpret_cast<uintptr_t>(src + w0) & 15) / sizeof(In));
                const int64_t head_end = min(w1, mis ? w0 + (V - mis) : w0);
                for (int64_t i = w0 + lane; i < head_end; i += TPR)
                    s0.take(load_as<In, Cmp>(src[i]), (uint32_t)(i - w0));
                const uint32_t hoff = (uint32_t)(head_end - w0);
                const uint32_t nvec = (uint32_t)((w1 - head_end) / V);
                const uint4 *v = reinterpret_cast<const uint4 *>(src + head_end);
                uint32_t i = lane;
                for (; i + 3 * TPR < nvec; i += 4 * TPR) {
                    const uint4 q0 = ldg_stream(v + i), q1 = ldg_stream(v + i + TPR), q2 = ldg_stream(v + i + 2 * TPR),
                                q3 = ldg_stream(v + i + 3 * TPR);
                    const In *e0 = reinterpret_cast<const In *>(&q0), *e1 = reinterpret_cast<const In *>(&q1),
                             *e2 = reinterpret_cast<const In *>(&q2), *e3 = reinterpret_cast<const In *>(&q3);
                    const uint32_t o = hoff + i * V;
                    Cmp c0[V], c1[V], c2[V], c3[V];
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        c0[k] = load_as<In, Cmp>(e0[k]);
                        c1[k] = load_as<In, Cmp>(e1[k]);
                        c2[k] = load_as<In, Cmp>(e2[k]);
                        c3[k] = load_as<In, Cmp>(e3[k]);
                    }
                    s0.take_vec(c0, o);
                    s1.take_vec(c1, o + TPR * V);
                    s2.take_vec(c2, o + 2 * TPR * V);
                    s3.take_vec(c3, o + 3 * TPR * V);
                }
                acc = Op::combine(acc, s0.pair(w0)); // head + stream 0 so far: all below the remainder in index order
                s0.init();
                for (; i < nvec; i += TPR) {
                    const uint4 q = ldg_stream(v + i);
                    const In *e = reinterpret_cast<const In *>(&q);
#pragma unroll
                    for (int k = 0; k < V; ++k)
                        s0.take(load_as<In, Cmp>(e[k]), hoff + i * V + k);
                }
                for (int64_t j = head_end + (int64_t)nvec * V + lane; j < w1; j += TPR)
                    s0.take(load_as<In, Cmp>(src[j]), (uint32_t)(j - w0));
                acc = Op::combine(acc, Op::combine(Op::combine(s0.pair(w0), s1.pair(w0)), Op::combine(s2.pair(w0), s3.pair(w0))));
            }
        } else {
            for (int64_t r = r0 + lane; r < r1; r += TPR) {
                const In *e = reinterpret_cast<const In *>(base + reduced_offset(p, r));
                acc = Op::combine(acc, Op::make(load_as<In, Cmp>(*e), r));
            }
        }
    }
    acc = group_reduce_pair<Cmp, Op, TPR>(acc);
    if (row_ok && lane == 0) {
        if (p.split > 1)
            reinterpret_cast<ArgPair<Cmp> *>(p.ws)[row * p.split + blockIdx.y] = acc;
        else
            reinterpret_cast<int64_t *>(p.out)[row] = acc.idx;
    }
}
template <typename Cmp, bool IS_MAX> __global__ void argreduce_finish_kernel(const __grid_constant__ Params p) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= p.n_out)
        return;
    typedef ArgOp<Cmp, IS_MAX> Op;
    const ArgPair<Cmp> *w = reinterpret_cast<const ArgPair<Cmp> *>(p.ws) + o * p.split;
    ArgPair<Cmp> a = Op::identity();
    for (int s = 0; s < p.split; ++s)
        a = Op::combine(a, w[s]);
    reinterpret_cast<int64_t *>(p.out)[o] = a.idx;
}

// ---------------------------------------------------------------------------------------------------
// host planning
// ---------------------------------------------------------------------------------------------------
namespace {

struct Group {
    std::vector<int64_t> shape, stride;
};
// merge adjacent dims (outer d, inner d+1) when stride[d] == stride[d+1] * shape[d+1]
void collapse(Group &g) {
    Group o;
    for (size_t d = 0; d < g.shape.size(); ++d) {
        if (g.shape[d] == 1)
            continue;
        if (!o.shape.empty() && o.stride.back() == g.stride[d] * g.shape[d]) {
            o.shape.back() *= g.shape[d];
            o.stride.back() = g.stride[d];
        } else {
            o.shape.push_back(g.shape[d]);
            o.stride.push_back(g.stride[d]);
        }
    }
    g = o;
}

axb_status fill_groups(const axb_tensor &in, const std::vector<bool> &reduced, Params &p) {
    Group k, r;
    for (int d = 0; d < in.ndim; ++d) {
        Group &g = reduced[d] ? r : k;
        g.shape.push_back(in.shape[d]);
        g.stride.push_back(in.strides[d]);
    }
    collapse(k);
    collapse(r);
    AXB_REQUIRE((int)k.shape.size() <= kMaxDims && (int)r.shape.size() <= kMaxDims, AXB_ERR_UNSUPPORTED,
                "axb_reduce: too many non-collapsible dims");
    p.nk = (int)k.shape.size();
    p.nr = (int)r.shape.size();
    p.n_out = 1;
    p.n_red = 1;
    for (int d = 0; d < p.nk; ++d) { // index 0 = innermost
        int s = p.nk - 1 - d;
        AXB_REQUIRE(k.shape[s] < (1ll << 31), AXB_ERR_UNSUPPORTED, "axb_reduce: dim too large");
        p.kshape[d] = (uint32_t)k.shape[s];
        p.kstride[d] = k.stride[s];
        FastDiv f = make_fastdiv((uint32_t)k.shape[s]);
        p.kdiv_m[d] = f.m;
        p.kdiv_s[d] = f.s;
        p.n_out *= k.shape[s];
    }
    for (int d = 0; d < p.nr; ++d) {
        int s = p.nr - 1 - d;
        AXB_REQUIRE(r.shape[s] < (1ll << 31), AXB_ERR_UNSUPPORTED, "axb_reduce: dim too large");
        p.rshape[d] = (uint32_t)r.shape[s];
        p.rstride[d] = r.stride[s];
        FastDiv f = make_fastdiv((uint32_t)r.shape[s]);
        p.rdiv_m[d] = f.m;
        p.rdiv_s[d] = f.s;
        p.n_red *= r.shape[s];
    }
    return AXB_OK;
}

enum Strategy { ROWS_CONTIG, ROWS_GENERIC, COLS };

struct Launch {
    Strategy strat;
    int tpr;  // ROWS: 32 or 256
    int vec;  // COLS
    dim3 grid;
};

// Picks the strategy, the split factor and the grid. `acc_size` sizes the workspace.
Launch choose(Params &p, size_t in_size, int sm_count, bool allow_cols) {
    Launch L{};
    const bool contig_r = p.nr == 0 || (p.nr == 1 && p.rstride[0] == (int64_t)in_size);
    const bool inner_kept_contig = p.nk >= 1 && p.kstride[0] == (int64_t)in_size && p.nr >= 1 &&
                                   std::abs(p.rstride[0]) >= (int64_t)in_size * p.kshape[0];
    const int64_t target_ctas = (int64_t)sm_count * 4;
    p.split = 1;
    p.chunk = p.n_red;
    if (allow_cols && !contig_r && inner_kept_contig && p.kshape[0] >= 16) {
        L.strat = COLS;
        int vec = (int)(16 / in_size);
        bool aligned = (reinterpret_cast<uintptr_t>(p.in) % 16) == 0 && p.kshape[0] % vec == 0;
        for (int d = 1; d < p.nk && aligned; ++d)
            aligned = p.kstride[d] % 16 == 0;
        for (int d = 0; d < p.nr && aligned; ++d)
            aligned = p.rstride[d] % 16 == 0;
        if (!aligned || (vec != 4 && vec != 8))
            vec = 1;
        L.vec = vec;
        int64_t gx = (p.n_out / vec + 31) / 32;
        if (gx < target_ctas && p.n_red >= 64) {
            int64_t want = std::min<int64_t>((target_ctas + gx - 1) / gx, p.n_red / 32);
            want = std::max<int64_t>(1, std::min<int64_t>(want, 1024));
            p.chunk = (p.n_red + want - 1) / want;
            p.split = (int)((p.n_red + p.chunk - 1) / p.chunk);
        }
        L.grid = dim3((unsigned)gx, (unsigned)p.split);
        return L;
    }
    L.strat = contig_r ? ROWS_CONTIG : ROWS_GENERIC;
    L.tpr = p.n_red >= 1024 ? 256 : 32;
    const int rows_per_cta = 256 / L.tpr;
    int64_t gx = (p.n_out + rows_per_cta - 1) / rows_per_cta;
    if (gx < target_ctas && p.n_red >= 16384) {
        int64_t want = std::min<int64_t>((target_ctas + gx - 1) / gx, p.n_red / 8192);
        want = std::max<int64_t>(1, std::min<int64_t>(want, 1024));
        p.chunk = (p.n_red + want - 1) / want;
        p.chunk = (p.chunk + 63) / 64 * 64; // keep chunk starts 16-byte aligned relative to the row start
        p.split = (int)((p.n_red + p.chunk - 1) / p.chunk);
        L.tpr = 256;
        gx = p.n_out;
    }
    L.grid = dim3((unsigned)gx, (unsigned)p.split);
    return L;
}

template <typename In, typename Acc, class Op>
axb_status run_typed(Params &p, DeviceCtx *ctx, cudaStream_t s) {
    Launch L = choose(p, sizeof(In), ctx->sm_count, true);
    void *ws = nullptr;
    if (p.split > 1 || p.peer_world > 1) {
        AXB_TRY(axb_alloc((size_t)p.n_out * p.split * sizeof(Acc), s, &ws));
        p.ws = static_cast<char *>(ws);
    }
    AXB_REQUIRE(L.grid.x > 0 && L.grid.y <= 65535, AXB_ERR_UNSUPPORTED, "axb_reduce: grid too large");
    switch (L.strat) {
    case ROWS_CONTIG:
        if (L.tpr == 256)
            reduce_rows_kernel<In, Acc, Op, 256, true><<<L.grid, 256, 0, s>>>(p);
        else
            reduce_rows_kernel<In, Acc, Op, 32, true><<<L.grid, 256, 0, s>>>(p);
        break;
    case ROWS_GENERIC:
        if (L.tpr == 256)
            reduce_rows_kernel<In, Acc, Op, 256, false><<<L.grid, 256, 0, s>>>(p);
        else
            reduce_rows_kernel<In, Acc, Op, 32, false><<<L.grid, 256, 0, s>>>(p);
        break;
    case COLS:
        if (sizeof(In) == 4 && L.vec == 4)
            reduce_cols_kernel<In, Acc, Op, (sizeof(In) == 4 ? 4 : 1)><<<L.grid, 256, 0, s>>>(p);
        else if (sizeof(In) == 2 && L.vec == 8)
            reduce_cols_kernel<In, Acc, Op, (sizeof(In) == 2 ? 8 : 1)><<<L.grid, 256, 0, s>>>(p);
        else
            reduce_cols_kernel<In, Acc, Op, 1><<<L.grid, 256, 0, s>>>(p);
        break;
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            if (ws)
                axb_free(ws, s);
            return fail(AXB_ERR_CUDA, "reduce kernel launch failed: %s", cudaGetErrorString(e));
        }
        count_launch();
    }
    if (p.split > 1 || p.peer_world > 1) {
        reduce_finish_kernel<Acc, Op><<<(unsigned)((p.n_out + 127) / 128), 128, 0, s>>>(p);
        cudaError_t e = cudaGetLastError();
        axb_free(ws, s); // stream-ordered: reusable only by later work on `s`
        if (e != cudaSuccess)
            return fail(AXB_ERR_CUDA, "reduce finish launch failed: %s", cudaGetErrorString(e));
        count_launch();
    }
    return AXB_OK;
}

template <typename In, typename Acc> axb_status run_op(int op, Params &p, DeviceCtx *ctx, cudaStream_t s) {
    switch (op) {
    case AXB_OP_SUM: case AXB_OP_MEAN: return run_typed<In, Acc, SumOp>(p, ctx, s);
    case AXB_OP_PROD: return run_typed<In, Acc, ProdOp>(p, ctx, s);
    case AXB_OP_MAX: return run_typed<In, Acc, MaxOp>(p, ctx, s);
    case AXB_OP_MIN: return run_typed<In, Acc, MinOp>(p, ctx, s);
    case AXB_OP_ANY: return run_typed<In, int32_t, AnyOp>(p, ctx, s);
    case AXB_OP_ALL: return run_typed<In, int32_t, AllOp>(p, ctx, s);
    default: return fail(AXB_ERR_UNSUPPORTED, "axb_reduce: op %d is not a reduction", op);
    }
}

template <typename In, typename Cmp, bool IS_MAX>
axb_status run_arg_typed(Params &p, DeviceCtx *ctx, cudaStream_t s) {
    Launch L = choose(p, sizeof(In), ctx->sm_count, false);
    void *ws = nullptr;
    if (p.split > 1) {
        AXB_TRY(axb_alloc((size_t)p.n_out * p.split * sizeof(ArgPair<Cmp>), s, &ws));
        p.ws = static_cast<char *>(ws);
    }
    AXB_REQUIRE(L.grid.x > 0 && L.grid.y <= 65535, AXB_ERR_UNSUPPORTED, "axb_argreduce: grid too large");
    if (L.strat == ROWS_CONTIG) {
        if (L.tpr == 256)
            argreduce_rows_kernel<In, Cmp, IS_MAX, 256, true><<<L.grid, 256, 0, s>>>(p);
        else
            argreduce_rows_kernel<In, Cmp, IS_MAX, 32, true><<<L.grid, 256, 0, s>>>(p);
    } else {
        if (L.tpr == 256)
            argreduce_rows_kernel<In, Cmp, IS_MAX, 256, false><<<L.grid, 256, 0, s>>>(p);
        else
            argreduce_rows_kernel<In, Cmp, IS_MAX, 32, false><<<L.grid, 256, 0, s>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        if (ws)
            axb_free(ws, s);
        return fail(AXB_ERR_CUDA, "argreduce kernel launch failed: %s", cudaGetErrorString(e));
    }
    count_launch();
    if (p.split > 1) {
        argreduce_finish_kernel<Cmp, IS_MAX><<<(unsigned)((p.n_out + 127) / 128), 128, 0, s>>>(p);
        e = cudaGetLastError();
        axb_free(ws, s);
        if (e != cudaSuccess)
            return fail(AXB_ERR_CUDA, "argreduce finish launch failed: %s", cudaGetErrorString(e));
        count_launch();
    }
    return AXB_OK;
}
template <typename In, typename Cmp> axb_status run_arg(bool is_max, Params &p, DeviceCtx *ctx, cudaStream_t s) {
    return is_max ? run_arg_typed<In, Cmp, true>(p, ctx, s) : run_arg_typed<In, Cmp, false>(p, ctx, s);
}

} // namespace
} // namespace red
} // namespace axb

using namespace axb;
using namespace axb::red;

// ---- sharded full reductions: peer mailboxes -------------------------------------------------------------------------
struct axb_peer_group {
    int rank = 0, world = 1;
    uint64_t *box[axb::red::kMaxPeers] = {};
    uint32_t seq = 0;          // calls issued so far; every rank issues the same sequence of sharded reductions
    int64_t global_count = 0;  // divisor of a sharded Mean (set per call)
    uint32_t *err = nullptr;   // pinned, device-visible: call counter of the first exchange that timed out (0 = none)
};
namespace axb {
int64_t &reduce_count_override() {
    static thread_local int64_t v = 0;
    return v;
}
} // namespace axb
static const axb_peer_group *&pending_peer_group() {
    static thread_local const axb_peer_group *g = nullptr;
    return g;
}

extern "C" {

// What the caller of the reference's CPU op observes (execute_reduction_typed keeps the input dtype; the graph
// executor widens a half-precision Mean to Float32, graph_executor.cpp:396-403).
int32_t axb_reduce_result_dtype(int32_t op, int32_t in_dtype) {
    if (op == AXB_OP_ARGMAX || op == AXB_OP_ARGMIN)
        return AXB_I64;
    if (op == AXB_OP_MEAN && is_half_dtype(in_dtype))
        return AXB_F32;
    return in_dtype;
}

axb_status axb_reduce(int32_t op, const axb_tensor *in, const int32_t *axes, int n_axes, int keep_dims,
                      const axb_tensor *out, axb_stream_t stream) {
    AXB_TRY(validate_tensor(in, "axb_reduce(in)"));
    AXB_TRY(validate_tensor(out, "axb_reduce(out)"));
    AXB_REQUIRE(n_axes >= 0 && n_axes <= in->ndim && (n_axes == 0 || axes), AXB_ERR_INVALID, "axb_reduce: bad axes");
    AXB_REQUIRE(!is_complex_dtype(in->dtype), AXB_ERR_UNSUPPORTED, "axb_reduce: complex dtypes have no GPU path");
    AXB_REQUIRE(is_contiguous(*out), AXB_ERR_UNSUPPORTED, "axb_reduce: out must be contiguous");
    std::vector<bool> reduced(in->ndim, n_axes == 0);
    for (int i = 0; i < n_axes; ++i) {
        AXB_REQUIRE(axes[i] >= 0 && axes[i] < in->ndim, AXB_ERR_INVALID,
                    "axb_reduce: axis %d out of range for rank %d (negative axes must be normalised by the caller)", axes[i],
                    in->ndim);
        AXB_REQUIRE(!reduced[axes[i]], AXB_ERR_INVALID, "axb_reduce: duplicate axis %d", axes[i]);
        reduced[axes[i]] = true;
    }
    const bool half_in = is_half_dtype(in->dtype);
    const bool od_ok = out->dtype == in->dtype || (op == AXB_OP_MEAN && half_in && out->dtype == AXB_F32) ||
                       ((op == AXB_OP_ANY || op == AXB_OP_ALL) && out->dtype == AXB_BOOL);
    AXB_REQUIRE(od_ok, AXB_ERR_INVALID, "axb_reduce: out dtype %d not valid for op %d on dtype %d", out->dtype, op, in->dtype);
    AXB_REQUIRE(!(in->dtype == AXB_BOOL && op != AXB_OP_ANY && op != AXB_OP_ALL && op != AXB_OP_MAX && op != AXB_OP_MIN),
                AXB_ERR_UNSUPPORTED, "axb_reduce: op %d on Bool is not supported", op);
    if (numel(*out) == 0)
        return AXB_OK;
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    Params p;
    memset(&p, 0, sizeof p);
    p.in = tensor_ptr(*in);
    p.out = tensor_ptr(*out);
    p.in_dtype = in->dtype;
    p.out_dtype = out->dtype;
    p.op = op;
    AXB_TRY(fill_groups(*in, reduced, p));
    // reduced shape per calculate_reduction_shape (cpu_operations.cpp:1280-1303): only the element count matters here
    AXB_REQUIRE(numel(*out) == p.n_out, AXB_ERR_SHAPE, "axb_reduce: out has %lld elements, the reduced shape has %lld",
                (long long)numel(*out), (long long)p.n_out);
    (void)keep_dims;
    p.count = p.n_red;
    if (int64_t &ov = axb::reduce_count_override(); ov > 0) {
        p.count = ov;
        ov = 0;
    }
    if (const axb_peer_group *g = pending_peer_group()) { // set by axb_reduce_sharded for exactly this call
        pending_peer_group() = nullptr;
        AXB_REQUIRE(p.n_out == 1 && (op == AXB_OP_SUM || op == AXB_OP_MEAN || op == AXB_OP_MAX || op == AXB_OP_MIN) &&
                        (in->dtype == AXB_F32 || in->dtype == AXB_BF16 || in->dtype == AXB_F16),
                    AXB_ERR_UNSUPPORTED, "axb_reduce_sharded: full sum / mean / max / min of f32, bf16 or f16 shards only");
        for (int r = 0; r < g->world; ++r)
            p.peer_box[r] = g->box[r];
        p.peer_rank = g->rank;
        p.peer_world = g->world;
        p.peer_seq = g->seq;
        p.peer_err = g->err;
        p.count = g->global_count;
    }
    switch (in->dtype) {
    case AXB_F32: return run_op<float, float>(op, p, ctx, s);
    case AXB_BF16: return run_op<bf16_t, float>(op, p, ctx, s);
    case AXB_F16: return run_op<f16_t, float>(op, p, ctx, s);
    case AXB_F64: return run_op<double, double>(op, p, ctx, s);
    case AXB_I32: return run_op<int32_t, int32_t>(op, p, ctx, s);
    case AXB_I64: return run_op<int64_t, int64_t>(op, p, ctx, s);
    case AXB_I16: return run_op<int16_t, int32_t>(op, p, ctx, s);
    case AXB_I8: return run_op<int8_t, int32_t>(op, p, ctx, s);
    case AXB_U8: case AXB_BOOL: return run_op<uint8_t, int32_t>(op, p, ctx, s);
    case AXB_U16: return run_op<uint16_t, int32_t>(op, p, ctx, s);
    case AXB_U32: return run_op<uint32_t, uint32_t>(op, p, ctx, s);
    case AXB_U64: return run_op<uint64_t, uint64_t>(op, p, ctx, s);
    default: return fail(AXB_ERR_UNSUPPORTED, "axb_reduce: dtype %d unsupported", in->dtype);
    }
}

axb_status axb_argreduce(int32_t op, const axb_tensor *in, int axis, int keep_dims, const axb_tensor *out,
                         axb_stream_t stream) {
    (void)keep_dims;
    AXB_TRY(validate_tensor(in, "axb_argreduce(in)"));
    AXB_TRY(validate_tensor(out, "axb_argreduce(out)"));
    AXB_REQUIRE(op == AXB_OP_ARGMAX || op == AXB_OP_ARGMIN, AXB_ERR_INVALID, "axb_argreduce: op must be ArgMax/ArgMin");
    AXB_REQUIRE(out->dtype == AXB_I64 && is_contiguous(*out), AXB_ERR_INVALID, "axb_argreduce: out must be contiguous Int64");
    AXB_REQUIRE(axis < in->ndim, AXB_ERR_INVALID, "axb_argreduce: axis %d out of range", axis);
    AXB_REQUIRE(!is_complex_dtype(in->dtype), AXB_ERR_UNSUPPORTED, "axb_argreduce: complex has no ordering");
    std::vector<bool> reduced(in->ndim, axis < 0);
    if (axis >= 0)
        reduced[axis] = true;
    else // flattened: the index is the C-order position, which needs a contiguous input (the reference flattens first)
        AXB_REQUIRE(is_contiguous(*in), AXB_ERR_UNSUPPORTED,
                    "axb_argreduce: flattened argmax needs a contiguous input (call axb_gather_strided first)");
    if (numel(*out) == 0)
        return AXB_OK;
    AXB_REQUIRE(numel(*in) > 0, AXB_ERR_INVALID, "axb_argreduce: empty input");
    cudaStream_t s;
    DeviceCtx *ctx;
    AXB_TRY(resolve_stream(stream, &s, &ctx));
    Params p;
    memset(&p, 0, sizeof p);
    p.in = tensor_ptr(*in);
    p.out = tensor_ptr(*out);
    p.in_dtype = in->dtype;
    p.out_dtype = AXB_I64;
    p.op = op;
    AXB_TRY(fill_groups(*in, reduced, p));
    AXB_REQUIRE(numel(*out) == p.n_out, AXB_ERR_SHAPE, "axb_argreduce: out has %lld elements, expected %lld",
                (long long)numel(*out), (long long)p.n_out);
    const bool mx = op == AXB_OP_ARGMAX;
    switch (in->dtype) {
    case AXB_F32: return run_arg<float, float>(mx, p, ctx, s);
    case AXB_BF16: return run_arg<bf16_t, float>(mx, p, ctx, s);
    case AXB_F16: return run_arg<f16_t, float>(mx, p, ctx, s);
    case AXB_F64: return run_arg<double, double>(mx, p, ctx, s);
    case AXB_I32: return run_arg<int32_t, int32_t>(mx, p, ctx, s);
    case AXB_I64: return run_arg<int64_t, int64_t>(mx, p, ctx, s);
    case AXB_I16: return run_arg<int16_t, int32_t>(mx, p, ctx, s);
    case AXB_I8: return run_arg<int8_t, int32_t>(mx, p, ctx, s);
    case AXB_U8: case AXB_BOOL: return run_arg<uint8_t, int32_t>(mx, p, ctx, s);
    case AXB_U16: return run_arg<uint16_t, int32_t>(mx, p, ctx, s);
    case AXB_U32: return run_arg<uint32_t, uint32_t>(mx, p, ctx, s);
    case AXB_U64: return run_arg<uint64_t, uint64_t>(mx, p, ctx, s);
    default: return fail(AXB_ERR_UNSUPPORTED, "axb_argreduce: dtype %d unsupported", in->dtype);
    }
}

// Mailbox = 2 banks x kMaxPeers 64-bit words in DEVICE memory of the calling rank's GPU, allocated with cudaMalloc (the
// caching pool sub-allocates, and an IPC handle names a whole allocation). The handle travels to the other ranks through
// the host plumbing (torch.distributed in axiom_b200/shard.py) and is opened there with axb_peer_mailbox_open.
axb_status axb_peer_mailbox_create(void **mailbox, unsigned char ipc_handle[64]) {
    AXB_REQUIRE(mailbox && ipc_handle, AXB_ERR_INVALID, "axb_peer_mailbox_create: null argument");
    cudaStream_t s;
    AXB_TRY(resolve_stream(nullptr, &s));
    void *ptr = nullptr;
    AXB_CUDA(cudaMalloc(&ptr, 2 * axb::red::kMaxPeers * sizeof(uint64_t)));
    AXB_CUDA(cudaMemset(ptr, 0, 2 * axb::red::kMaxPeers * sizeof(uint64_t)));
    AXB_CUDA(cudaDeviceSynchronize());
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    AXB_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(ipc_handle, &h, 64);
    *mailbox = ptr;
    return AXB_OK;
}
axb_status axb_peer_mailbox_open(const unsigned char ipc_handle[64], void **peer_mailbox) {
    AXB_REQUIRE(ipc_handle && peer_mailbox, AXB_ERR_INVALID, "axb_peer_mailbox_open: null argument");
    cudaStream_t s;
    AXB_TRY(resolve_stream(nullptr, &s));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    AXB_CUDA(cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess));
    return AXB_OK;
}
axb_status axb_peer_mailbox_close(void *peer_mailbox) {
    AXB_CUDA(cudaIpcCloseMemHandle(peer_mailbox));
    return AXB_OK;
}
axb_status axb_peer_mailbox_destroy(void *mailbox) {
    AXB_CUDA(cudaFree(mailbox));
    return AXB_OK;
}
axb_status axb_peer_group_create(int rank, int world, void *const *mailboxes, axb_peer_group **group) {
    AXB_REQUIRE(group && mailboxes && world >= 1 && world <= axb::red::kMaxPeers && rank >= 0 && rank < world, AXB_ERR_INVALID,
                "axb_peer_group_create: need 1 <= world <= %d and 0 <= rank < world", axb::red::kMaxPeers);
    axb_peer_group *g = new axb_peer_group();
    g->rank = rank;
    g->world = world;
    for (int r = 0; r < world; ++r) {
        if (!mailboxes[r]) {
            delete g;
            return fail(AXB_ERR_INVALID, "axb_peer_group_create: mailbox %d is null", r);
        }
        g->box[r] = static_cast<uint64_t *>(mailboxes[r]);
    }
    void *e = nullptr;
    if (cudaHostAlloc(&e, sizeof(uint32_t), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        delete g;
        return fail(AXB_ERR_CUDA, "axb_peer_group_create: cannot allocate the error word");
    }
    g->err = static_cast<uint32_t *>(e);
    *g->err = 0;
    *group = g;
    return AXB_OK;
}
axb_status axb_peer_group_status(const axb_peer_group *group) {
    AXB_REQUIRE(group, AXB_ERR_INVALID, "axb_peer_group_status: null group");
    const uint32_t bad = group->err ? *(volatile uint32_t *)group->err : 0;
    if (bad)
        return fail(AXB_ERR_CUDA, "axb_reduce_sharded: exchange %u timed out on rank %d (a peer never delivered its partial); "
                                  "the result of that call is NaN", bad, group->rank);
    return AXB_OK;
}
axb_status axb_peer_group_destroy(axb_peer_group *group) {
    if (group && group->err)
        cudaFreeHost(group->err);
    delete group;
    return AXB_OK;
}
axb_status axb_reduce_sharded(int32_t op, const axb_tensor *in_shard, int64_t global_count, axb_peer_group *group,
                              const axb_tensor *out, axb_stream_t stream) {
    AXB_REQUIRE(group, AXB_ERR_INVALID, "axb_reduce_sharded: null group");
    AXB_REQUIRE(in_shard && numel(*in_shard) > 0, AXB_ERR_INVALID, "axb_reduce_sharded: every rank must own a non-empty shard");
    if (group->world == 1)
        return axb_reduce(op, in_shard, nullptr, 0, 0, out, stream);
    AXB_TRY(axb_peer_group_status(group)); // an earlier exchange timed out: every later call fails loudly
    // everything that can be rejected is checked BEFORE the call counter moves: a rank that fails validation must not
    // leave its sequence one ahead of the others (they would time out on every later call)
    AXB_TRY(validate_tensor(in_shard, "axb_reduce_sharded(in)"));
    AXB_TRY(validate_tensor(out, "axb_reduce_sharded(out)"));
    AXB_REQUIRE((op == AXB_OP_SUM || op == AXB_OP_MEAN || op == AXB_OP_MAX || op == AXB_OP_MIN) &&
                    (in_shard->dtype == AXB_F32 || in_shard->dtype == AXB_BF16 || in_shard->dtype == AXB_F16),
                AXB_ERR_UNSUPPORTED, "axb_reduce_sharded: full sum / mean / max / min of f32, bf16 or f16 shards only");
    AXB_REQUIRE(numel(*out) == 1 && is_contiguous(*out) &&
                    (out->dtype == in_shard->dtype || out->dtype == axb_reduce_result_dtype(op, in_shard->dtype)),
                AXB_ERR_SHAPE, "axb_reduce_sharded: out must be one element of the reduction's result dtype");
    AXB_REQUIRE(in_shard->ndim <= AXB_MAX_DIMS && global_count >= numel(*in_shard), AXB_ERR_INVALID,
                "axb_reduce_sharded: global_count is smaller than the shard");
    const uint32_t seq_before = group->seq;
    ++group->seq;
    if (group->seq == 0) // 0 is the mailbox's initial tag
        ++group->seq;
    group->global_count = global_count;
    pending_peer_group() = group;
    const axb_status st = axb_reduce(op, in_shard, nullptr, 0, 0, out, stream);
    pending_peer_group() = nullptr;
    if (st != AXB_OK)
        group->seq = seq_before; // nothing was enqueued: the exchange did not happen on this rank
    return st;
}

} // extern "C"
