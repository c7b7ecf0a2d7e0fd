// reduce_common.cuh — shared device helpers of kernel family 2 (reductions, argmax, softmax, norms).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "axb.h"

namespace axb {
namespace red {

constexpr int kMaxDims = 8;

struct bf16_t {
    uint16_t bits;
};
struct f16_t {
    uint16_t bits;
};

constexpr int kMaxPeers = 16;
struct Params {
    const char *in;
    char *out;
    char *ws;          // partials of split reductions
    int32_t nk, nr;    // collapsed kept / reduced dims; index 0 = innermost
    uint32_t kshape[kMaxDims], rshape[kMaxDims];
    uint32_t kdiv_m[kMaxDims], kdiv_s[kMaxDims], rdiv_m[kMaxDims], rdiv_s[kMaxDims];
    int64_t kstride[kMaxDims], rstride[kMaxDims]; // bytes
    int64_t n_out, n_red;
    int64_t count;     // divisor of Mean
    int64_t chunk;     // reduced elements per split
    int32_t split;
    int32_t op, in_dtype, out_dtype;
    // sharded FULL reduction (axb_reduce_sharded): the finish kernel also exchanges the shard's partial with the other
    // ranks through their mailboxes in peer (NVLink / NVSwitch) memory — see peer exchange in reduce.cu
    uint64_t *peer_box[kMaxPeers]; // mailbox of every rank (own included), mapped into this process
    int32_t peer_rank, peer_world; // peer_world <= 1: not sharded
    uint32_t peer_seq;             // call counter, identical on all ranks; selects the bank and tags every word
    uint32_t *peer_err;            // host-mapped error word of the group: a timed-out exchange stores its call counter here
};

__device__ __forceinline__ uint32_t fastdiv32(uint32_t n, uint32_t m, uint32_t s) { return (__umulhi(n, m) + n) >> s; }

// byte offset of output index o (C order over the kept dims)
__device__ __forceinline__ int64_t kept_offset(const Params &p, int64_t o) {
    int64_t off = 0;
#pragma unroll 1
    for (int d = 0; d < p.nk; ++d) {
        if (d == p.nk - 1) {
            off += o * p.kstride[d];
        } else {
            int64_t q = o / p.kshape[d];
            off += (o - q * p.kshape[d]) * p.kstride[d];
            o = q;
        }
    }
    return off;
}
__device__ __forceinline__ int64_t reduced_offset(const Params &p, int64_t r) {
    if (p.nr == 1)
        return r * p.rstride[0];
    int64_t off = 0;
#pragma unroll 1
    for (int d = 0; d < p.nr; ++d) {
        if (d == p.nr - 1) {
            off += r * p.rstride[d];
        } else {
            int64_t q = r / p.rshape[d];
            off += (r - q * p.rshape[d]) * p.rstride[d];
            r = q;
        }
    }
    return off;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4 *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ float h2f(uint16_t h) {
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ uint16_t f2h(float f) {
    uint16_t h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f));
    return h;
}
// reference bfloat16_t(float): NaN -> 0x7FC0, else round-to-nearest-even (include/axiom/bfloat16.hpp:16-27)
__device__ __forceinline__ uint16_t f2bf(float f) {
    if (f != f)
        return 0x7FC0u;
    uint32_t b = __float_as_uint(f);
    b += 0x7FFFu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}
__device__ __forceinline__ float bf2f(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// Exact three-way bf16 split of TWO floats with the hardware converter: x = hi + mid + lo, each piece packed as
// (b in the upper half, a in the lower half). 11 instructions per pair instead of ~40 through the software rounding.
// NaN / inf inputs give NaN pieces either way (inf - inf), so the NaN encoding of the converter does not matter here.
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t &hi, uint32_t &mid, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float a1 = a - __uint_as_float(hi << 16), b1 = b - __uint_as_float(hi & 0xFFFF0000u); // exact
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(b1), "f"(a1));
    const float a2 = a1 - __uint_as_float(mid << 16), b2 = b1 - __uint_as_float(mid & 0xFFFF0000u); // exact
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b2), "f"(a2));
}

// storage element -> compute type
template <typename In, typename T> __device__ __forceinline__ T load_as(In v) { return (T)v; }
template <> __device__ __forceinline__ float load_as<bf16_t, float>(bf16_t v) { return bf2f(v.bits); }
template <> __device__ __forceinline__ float load_as<f16_t, float>(f16_t v) { return h2f(v.bits); }

// ---- reduction operators --------------------------------------------------------------------------
struct SumOp {
    template <typename A> __device__ static A identity() { return A(0); }
    template <typename In, typename A> __device__ static A load(In v) { return load_as<In, A>(v); }
    template <typename A> __device__ static A combine(A a, A b) { return a + b; }
};
struct ProdOp {
    template <typename A> __device__ static A identity() { return A(1); }
    template <typename In, typename A> __device__ static A load(In v) { return load_as<In, A>(v); }
    template <typename A> __device__ static A combine(A a, A b) { return a * b; }
};
template <typename A> struct Limits;
template <> struct Limits<float> {
    __device__ static float lowest() { return -3.402823466e+38f; }
    __device__ static float highest() { return 3.402823466e+38f; }
};
template <> struct Limits<double> {
    __device__ static double lowest() { return -1.7976931348623157e+308; }
    __device__ static double highest() { return 1.7976931348623157e+308; }
};
template <> struct Limits<int32_t> {
    __device__ static int32_t lowest() { return INT32_MIN; }
    __device__ static int32_t highest() { return INT32_MAX; }
};
template <> struct Limits<uint32_t> {
    __device__ static uint32_t lowest() { return 0u; }
    __device__ static uint32_t highest() { return 0xFFFFFFFFu; }
};
template <> struct Limits<int64_t> {
    __device__ static int64_t lowest() { return INT64_MIN; }
    __device__ static int64_t highest() { return INT64_MAX; }
};
template <> struct Limits<uint64_t> {
    __device__ static uint64_t lowest() { return 0ull; }
    __device__ static uint64_t highest() { return 0xFFFFFFFFFFFFFFFFull; }
};
// MaxFunc/MinFunc identity = lowest()/max() (cpu_operations.hpp), std::max/min combine
struct MaxOp {
    template <typename A> __device__ static A identity() { return Limits<A>::lowest(); }
    template <typename In, typename A> __device__ static A load(In v) { return load_as<In, A>(v); }
    template <typename A> __device__ static A combine(A a, A b) { return (a < b) ? b : a; }
};
struct MinOp {
    template <typename A> __device__ static A identity() { return Limits<A>::highest(); }
    template <typename In, typename A> __device__ static A load(In v) { return load_as<In, A>(v); }
    template <typename A> __device__ static A combine(A a, A b) { return (b < a) ? b : a; }
};
template <typename In> __device__ __forceinline__ bool nonzero(In v) { return v != In(0); }
template <> __device__ __forceinline__ bool nonzero<bf16_t>(bf16_t v) { return bf2f(v.bits) != 0.0f; }
template <> __device__ __forceinline__ bool nonzero<f16_t>(f16_t v) { return h2f(v.bits) != 0.0f; }
struct AnyOp {
    template <typename A> __device__ static A identity() { return A(0); }
    template <typename In, typename A> __device__ static A load(In v) { return nonzero<In>(v) ? A(1) : A(0); }
    template <typename A> __device__ static A combine(A a, A b) { return a | b; }
};
struct AllOp {
    template <typename A> __device__ static A identity() { return A(1); }
    template <typename In, typename A> __device__ static A load(In v) { return nonzero<In>(v) ? A(1) : A(0); }
    template <typename A> __device__ static A combine(A a, A b) { return a & b; }
};

template <typename In, typename Acc, class Op> __device__ __forceinline__ Acc fold_vec(Acc acc, uint4 q) {
    constexpr int V = 16 / sizeof(In);
    const In *w = reinterpret_cast<const In *>(&q);
    // pairwise inside the vector keeps the dependent chain short
    Acc t[V];
#pragma unroll
    for (int k = 0; k < V; ++k)
        t[k] = Op::template load<In, Acc>(w[k]);
#pragma unroll
    for (int s = V / 2; s >= 1; s /= 2)
#pragma unroll
        for (int k = 0; k < s; ++k)
            t[k] = Op::combine(t[k], t[k + s]);
    return Op::combine(acc, t[0]);
}

template <typename T> __device__ __forceinline__ T shfl_down_t(T v, int d) { return __shfl_down_sync(0xFFFFFFFFu, v, d); }
template <> __device__ __forceinline__ int64_t shfl_down_t<int64_t>(int64_t v, int d) {
    return (int64_t)__shfl_down_sync(0xFFFFFFFFu, (long long)v, d);
}
template <> __device__ __forceinline__ uint64_t shfl_down_t<uint64_t>(uint64_t v, int d) {
    return (uint64_t)__shfl_down_sync(0xFFFFFFFFu, (unsigned long long)v, d);
}

// warp-shuffle, then (TPR == 256) shared-memory combine of the 8 warp results; valid in lane 0 of the group
template <typename Acc, class Op, int TPR> __device__ __forceinline__ Acc group_reduce(Acc a) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        a = Op::combine(a, shfl_down_t<Acc>(a, d));
    if (TPR == 256) {
        __shared__ Acc warp_part[8];
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0)
            warp_part[w] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 1; k < 8; ++k)
                a = Op::combine(a, warp_part[k]);
        }
    }
    return a;
}

// acc -> output element: mean scaling, rounding to the input's storage type, conversion to the output dtype
template <typename Acc, class Op> __device__ __forceinline__ void finalize(const Params &p, int64_t o, Acc a) {
    if (p.op == AXB_OP_MEAN && p.count != 0) {
        // sum / N in the accumulation type; integer mean truncates (C++ integer division). Small ints wrapped first.
        if (p.in_dtype == AXB_I8)
            a = (Acc)(int8_t)a;
        else if (p.in_dtype == AXB_U8)
            a = (Acc)(uint8_t)a;
        else if (p.in_dtype == AXB_I16)
            a = (Acc)(int16_t)a;
        else if (p.in_dtype == AXB_U16)
            a = (Acc)(uint16_t)a;
        a = (Acc)(a / (Acc)p.count);
    }
    switch (p.out_dtype) {
    case AXB_F32: {
        float f = (float)a;
        if (p.in_dtype == AXB_BF16)
            f = bf2f(f2bf(f)); // half Mean: rounded to the half type first, then widened (graph_executor.cpp:396-403)
        else if (p.in_dtype == AXB_F16)
            f = h2f(f2h(f));
        reinterpret_cast<float *>(p.out)[o] = f;
        break;
    }
    case AXB_BF16: reinterpret_cast<uint16_t *>(p.out)[o] = f2bf((float)a); break;
    case AXB_F16: reinterpret_cast<uint16_t *>(p.out)[o] = f2h((float)a); break;
    case AXB_F64: reinterpret_cast<double *>(p.out)[o] = (double)a; break;
    case AXB_I32: case AXB_U32: reinterpret_cast<uint32_t *>(p.out)[o] = (uint32_t)a; break;
    case AXB_I64: case AXB_U64: reinterpret_cast<uint64_t *>(p.out)[o] = (uint64_t)a; break;
    case AXB_I16: case AXB_U16: reinterpret_cast<uint16_t *>(p.out)[o] = (uint16_t)a; break;
    case AXB_I8: case AXB_U8: reinterpret_cast<uint8_t *>(p.out)[o] = (uint8_t)a; break;
    case AXB_BOOL: reinterpret_cast<uint8_t *>(p.out)[o] = (a != Acc(0)) ? 1 : 0; break;
    default: break;
    }
}

// ---- argmax / argmin ------------------------------------------------------------------------------
template <typename C> struct ArgPair {
    C val;
    int64_t idx;
};
// A candidate only exists if it strictly beats the reference's initial best (lowest() / max()); everything else
// (NaN, -inf for argmax, ...) maps to the identity (init, 0). Ties keep the smaller index == first wins.
template <typename C, bool IS_MAX> struct ArgOp {
    __device__ static ArgPair<C> identity() { return {IS_MAX ? Limits<C>::lowest() : Limits<C>::highest(), 0}; }
    __device__ static ArgPair<C> make(C v, int64_t i) {
        const bool cand = IS_MAX ? (v > Limits<C>::lowest()) : (v < Limits<C>::highest());
        return cand ? ArgPair<C>{v, i} : identity();
    }
    __device__ static ArgPair<C> combine(ArgPair<C> a, ArgPair<C> b) {
        const bool take_b = IS_MAX ? (b.val > a.val || (b.val == a.val && b.idx < a.idx))
                                   : (b.val < a.val || (b.val == a.val && b.idx < a.idx));
        return take_b ? b : a;
    }
};
// Per-thread scan over elements presented in increasing index order (32-bit offsets relative to a window start).
template <typename C, bool IS_MAX> struct ArgScan {
    C best;
    uint32_t off;
    __device__ __forceinline__ void init() {
        best = IS_MAX ? Limits<C>::lowest() : Limits<C>::highest();
        off = 0;
    }
    __device__ __forceinline__ void take(C v, uint32_t o) {
        const bool t = IS_MAX ? (v > best) : (v < best);
        best = t ? v : best;
        off = t ? o : off;
    }
    // V consecutive elements starting at offset `base`: the value chain is the same, but only a lane-local position
    // (a select between immediates) is tracked per element; the offset is formed once per vector. `best` only ever
    // improves, so the LAST improving element of the vector is the vector's first extremum.
    template <int V> __device__ __forceinline__ void take_vec(const C (&v)[V], uint32_t base) {
        int kk = -1;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const bool t = IS_MAX ? (v[k] > best) : (v[k] < best);
            best = t ? v[k] : best;
            kk = t ? k : kk;
        }
        off = kk >= 0 ? base + (uint32_t)kk : off;
    }
    __device__ __forceinline__ ArgPair<C> pair(int64_t window_start) const {
        const bool hit = IS_MAX ? (best > Limits<C>::lowest()) : (best < Limits<C>::highest());
        return hit ? ArgPair<C>{best, window_start + (int64_t)off} : ArgOp<C, IS_MAX>::identity();
    }
};
template <typename C, class Op, int TPR> __device__ __forceinline__ ArgPair<C> group_reduce_pair(ArgPair<C> a) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        ArgPair<C> b;
        b.val = shfl_down_t<C>(a.val, d);
        b.idx = shfl_down_t<int64_t>(a.idx, d);
        a = Op::combine(a, b);
    }
    if (TPR == 256) {
        __shared__ ArgPair<C> warp_part[8];
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0)
            warp_part[w] = a;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 1; k < 8; ++k)
                a = Op::combine(a, warp_part[k]);
        }
    }
    return a;
}

} // namespace red
} // namespace axb
