// cuda_plan.cpp — see cuda_plan.hpp: record / replay of the C-ABI call sequence of one graph materialisation.
#include "cuda_plan.hpp"

#include <cstring>
#include <mutex>
#include <unordered_map>
#include <unordered_set>

#include "axiom/graph/graph_node.hpp"
#include "axiom/graph/graph_signature.hpp"

namespace axiom {
namespace backends {
namespace cuda {
namespace plan {

namespace {

using graph::GraphNode;

enum Fn : int32_t {
    F_ELEMENTWISE, F_ELEMENTWISE_REDUCE, F_MATMUL, F_SOFTMAX, F_LAYER_NORM, F_RMS_NORM, F_REDUCE, F_ARGREDUCE, F_CAST,
    F_GATHER_STRIDED, F_IM2COL, F_BINARY, F_UNARY, F_WHERE, F_MASKED_FILL, F_PAD, F_CONV2D, F_FILL, F_POOL, F_SORT, F_ELEMENTWISE_REDUCE_AXIS, F_ELEMENTWISE_SOFTMAX, F_ELEMENTWISE_NORM
};
enum Kind : int8_t { K_NONE = 0, K_LEAF = 1, K_TEMP = 2 };

struct TRef {
    int8_t kind = K_NONE;
    int32_t idx = -1;
    axb_tensor d{}; // base = nullptr; byte_offset relative to the slot's base pointer
};
constexpr int kMaxT = 10;
struct Call {
    Fn fn;
    TRef t[kMaxT];
    std::vector<axb_instr> prog;
    int64_t iv[20] = {};
    double dv[2] = {};
};
struct Temp {
    size_t bytes = 0;
    int last = -1;   // index of the last call that touches it (-1: none)
    int up_leaf = -1; // uploaded from this (CPU) leaf ...
    size_t up_off = 0, up_bytes = 0;
};
struct Plan {
    bool cacheable = false;
    std::vector<Call> calls;
    std::vector<Temp> temps;
    TRef result;
    Shape shape;
    Strides strides;
    DType dtype = DType::Float32;
    size_t n_leaves = 0;
    uint64_t tick = 0;
};
struct Leaf {
    std::shared_ptr<Storage> st;
    bool gpu = false;
    const char *ptr = nullptr; // device pointer (GPU) or host pointer (CPU)
    size_t bytes = 0;
};

bool is_leaf(const GraphNode *n) { return n->is_constant || n->is_materialized_; }

inline uint64_t mix(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h * 0xff51afd7ed558ccdull;
}

// Leaves of the lazy graph under `root` in deterministic DFS order, and the layout part of the key.
bool collect(const GraphNode *root, std::vector<Leaf> *leaves, uint64_t *layout_hash) {
    std::unordered_set<const GraphNode *> seen;
    std::vector<const GraphNode *> stk{root};
    std::unordered_map<const void *, uint32_t> first_of; // storage pointer -> first leaf that uses it (alias structure)
    uint64_t h = 0x1234567ull;
    while (!stk.empty()) {
        const GraphNode *n = stk.back();
        stk.pop_back();
        if (!n || !seen.insert(n).second)
            continue;
        if (is_leaf(n) && n != root) {
            std::shared_ptr<Storage> st = n->is_constant ? n->constant_storage : n->cached_result_;
            const Strides &sd = n->is_constant ? n->constant_strides : n->cached_strides_;
            const Shape &sh = n->is_constant ? n->output_shape : n->cached_shape_;
            if (!st) { // an absent optional operand (no bias): part of the structure, not a tensor
                h = mix(h, 0xABu);
                continue;
            }
            Leaf lf;
            lf.st = st;
            lf.gpu = st->device() == Device::GPU;
            lf.bytes = st->size_bytes();
            if (lf.gpu) {
                auto *cs = dynamic_cast<CudaStorage *>(st.get());
                if (!cs)
                    return false;
                lf.ptr = static_cast<const char *>(cs->device_ptr());
            } else {
                lf.ptr = lf.bytes ? static_cast<const char *>(st->data()) : nullptr;
            }
            const size_t off = n->is_constant ? n->constant_offset : 0;
            const uint32_t me = (uint32_t)leaves->size();
            auto ins = first_of.emplace(static_cast<const void *>(lf.ptr), me);
            h = mix(h, lf.gpu ? 1 : 2);
            h = mix(h, (uint64_t)n->output_dtype);
            h = mix(h, ins.first->second); // alias class
            h = mix(h, off);
            h = mix(h, (uint64_t)(reinterpret_cast<uintptr_t>(lf.ptr) + off) & 15u);
            h = mix(h, lf.bytes);
            h = mix(h, sh.size());
            for (auto d : sh)
                h = mix(h, d);
            for (auto s : sd)
                h = mix(h, (uint64_t)s);
            leaves->push_back(std::move(lf));
            continue;
        }
        // inputs pushed in reverse so they pop in order
        for (size_t i = n->inputs.size(); i-- > 0;)
            stk.push_back(n->inputs[i].get());
    }
    *layout_hash = h;
    return true;
}

struct Rec {
    bool active = false, bad = false;
    uint64_t key = 0;
    std::vector<Leaf> leaves;
    std::unordered_map<const void *, std::pair<int8_t, int32_t>> slot; // base pointer -> (kind, index)
    Plan plan;
    uint64_t launches0 = 0, rec_launches = 0, rec_h2d = 0;
    axb_transfer_stats x0{};
    const char *why = "";
};
thread_local Rec g_rec;
// key + leaves of the last lookup on this thread (try_replay -> begin_record without recomputing)
thread_local struct {
    bool valid = false, known_uncacheable = false;
    const GraphNode *root = nullptr;
    uint64_t key = 0;
    std::vector<Leaf> leaves;
} g_pending;

std::mutex g_mu;
std::unordered_map<uint64_t, std::shared_ptr<Plan>> g_cache;
Stats g_stats;
uint64_t g_tick = 0;
constexpr size_t kMaxPlans = 512; // GraphCache::MAX_SIZE (include/axiom/graph/graph_cache.hpp:35)

TRef resolve(const axb_tensor *t) {
    TRef r;
    if (!t)
        return r;
    r.d = *t;
    auto it = g_rec.slot.find(t->base);
    if (it == g_rec.slot.end()) {
        if (t->base) { // memory the recorder has never seen (neither a leaf nor a temporary of this lowering)
            g_rec.bad = true;
            g_rec.why = "tensor outside the recorded graph";
        }
        r.d.base = nullptr;
        return r;
    }
    r.kind = it->second.first;
    r.idx = it->second.second;
    r.d.base = nullptr;
    return r;
}

// call bookkeeping shared by every wrapper
struct Scope {
    uint64_t l0;
    bool on;
    Scope() : l0(0), on(g_rec.active && !g_rec.bad) {
        if (on)
            l0 = axb_launch_count();
    }
    Call *open(Fn fn) {
        g_rec.rec_launches += axb_launch_count() - l0;
        g_rec.plan.calls.emplace_back();
        Call &c = g_rec.plan.calls.back();
        c.fn = fn;
        return &c;
    }
};

axb_status invoke(const Call &c, axb_tensor *T) {
    auto tp = [&](int i) -> const axb_tensor * { return c.t[i].kind == K_NONE ? nullptr : &T[i]; };
    switch (c.fn) {
    case F_ELEMENTWISE: {
        const int n_in = (int)c.iv[0];
        return axb_elementwise(c.prog.data(), (int)c.prog.size(), T, n_in, &T[n_in], nullptr);
    }
    case F_ELEMENTWISE_REDUCE: {
        const int n_in = (int)c.iv[0];
        return axb_elementwise_reduce(c.prog.data(), (int)c.prog.size(), T, n_in, (int32_t)c.iv[1], &T[n_in], nullptr);
    }
    case F_ELEMENTWISE_REDUCE_AXIS: {
        const int n_in = (int)c.iv[0];
        return axb_elementwise_reduce_axis(c.prog.data(), (int)c.prog.size(), T, n_in, (int32_t)c.iv[1], (int)c.iv[2], (int)c.iv[3], &T[n_in],
                                           nullptr);
    }
    case F_ELEMENTWISE_SOFTMAX: {
        const int n_in = (int)c.iv[0];
        return axb_elementwise_softmax(c.prog.data(), (int)c.prog.size(), T, n_in, (int)c.iv[1], (int32_t)c.iv[2], &T[n_in], nullptr);
    }
    case F_ELEMENTWISE_NORM: { // T: inputs..., out, weight, bias
        const int n_in = (int)c.iv[0];
        return axb_elementwise_norm(c.prog.data(), (int)c.prog.size(), T, n_in, (int32_t)c.iv[1], tp(n_in + 1), tp(n_in + 2), (float)c.dv[0],
                                    &T[n_in], nullptr);
    }
    case F_MATMUL: {
        axb_matmul_epilogue e{};
        e.bias = tp(2);
        e.activation = (int32_t)c.iv[2];
        e.flags = (int32_t)c.iv[3];
        return axb_matmul(&T[0], &T[1], (int)c.iv[0], (int)c.iv[1], c.iv[4] ? &e : nullptr, &T[3], nullptr);
    }
    case F_SOFTMAX: return axb_softmax(&T[0], (int)c.iv[0], (int)c.iv[1], &T[1], nullptr);
    case F_LAYER_NORM: return axb_layer_norm(&T[0], tp(1), tp(2), (int)c.iv[0], (float)c.dv[0], &T[3], nullptr);
    case F_RMS_NORM: return axb_rms_norm(&T[0], tp(1), (int)c.iv[0], (float)c.dv[0], &T[2], nullptr);
    case F_REDUCE: {
        int32_t axes[AXB_MAX_DIMS];
        const int na = (int)c.iv[1];
        for (int i = 0; i < na; ++i)
            axes[i] = (int32_t)c.iv[3 + i];
        return axb_reduce((int32_t)c.iv[0], &T[0], axes, na, (int)c.iv[2], &T[1], nullptr);
    }
    case F_ARGREDUCE: return axb_argreduce((int32_t)c.iv[0], &T[0], (int)c.iv[1], (int)c.iv[2], &T[1], nullptr);
    case F_CAST: return axb_cast(&T[0], &T[1], nullptr);
    case F_GATHER_STRIDED: return axb_gather_strided(&T[0], &T[1], nullptr);
    case F_IM2COL: return axb_im2col_1d(&T[0], (int)c.iv[0], (int)c.iv[1], (int)c.iv[2], (int)c.iv[3], &T[1], nullptr);
    case F_BINARY: return axb_binary((int32_t)c.iv[0], &T[0], &T[1], &T[2], nullptr);
    case F_UNARY: return axb_unary((int32_t)c.iv[0], &T[0], &T[1], c.dv[0], nullptr);
    case F_WHERE: return axb_where(&T[0], &T[1], &T[2], &T[3], nullptr);
    case F_MASKED_FILL: return axb_masked_fill(&T[0], &T[1], c.dv[0], &T[2], nullptr);
    case F_PAD: return axb_pad_constant(&T[0], &c.iv[0], &c.iv[8], c.dv[0], &T[1], nullptr);
    case F_CONV2D: {
        const int32_t st[2] = {(int32_t)c.iv[0], (int32_t)c.iv[1]}, pd[2] = {(int32_t)c.iv[2], (int32_t)c.iv[3]},
                      dl[2] = {(int32_t)c.iv[4], (int32_t)c.iv[5]};
        return axb_conv(&T[0], &T[1], tp(2), (int)c.iv[7], st, pd, dl, (int)c.iv[6], (int)c.iv[8], &T[3], nullptr);
    }
    case F_FILL: return axb_fill(&T[0], c.dv[0], nullptr);
    case F_POOL: {
        int32_t k[3], st[3], pd[3];
        for (int d = 0; d < 3; ++d) {
            k[d] = (int32_t)c.iv[3 + d];
            st[d] = (int32_t)c.iv[6 + d];
            pd[d] = (int32_t)c.iv[9 + d];
        }
        return axb_pool(&T[0], (int)c.iv[0], (int)c.iv[1], k, st, pd, (int)c.iv[2], &T[1], nullptr);
    }
    case F_SORT: return axb_sort(&T[0], (int)c.iv[0], (int)c.iv[1], tp(1), tp(2), nullptr);
    }
    return AXB_ERR_INVALID;
}

} // namespace

bool enabled() {
    static const bool off = std::getenv("AXIOM_CUDA_NO_PLAN_CACHE") != nullptr; // A/B switch
    return !off;
}
Stats stats() {
    std::lock_guard<std::mutex> lk(g_mu);
    Stats s = g_stats;
    s.entries = g_cache.size();
    return s;
}
void clear() {
    std::lock_guard<std::mutex> lk(g_mu);
    g_cache.clear();
    g_stats = Stats{};
}

bool recording() { return g_rec.active; }
void note_alloc(void *p, size_t bytes) {
    if (!g_rec.active || g_rec.bad || !p)
        return;
    Temp t;
    t.bytes = bytes;
    g_rec.plan.temps.push_back(t);
    g_rec.slot[p] = {K_TEMP, (int32_t)g_rec.plan.temps.size() - 1}; // a recycled address now names the NEW temporary
}
void note_upload(void *dst, const void *src, size_t n) {
    if (!g_rec.active || g_rec.bad)
        return;
    auto it = g_rec.slot.find(dst);
    if (it == g_rec.slot.end() || it->second.first != K_TEMP) {
        g_rec.bad = true;
        g_rec.why = "upload into memory that is not a temporary of this lowering";
        return;
    }
    const char *s = static_cast<const char *>(src);
    for (size_t k = 0; k < g_rec.leaves.size(); ++k) {
        const Leaf &lf = g_rec.leaves[k];
        if (!lf.gpu && lf.ptr && s >= lf.ptr && s + n <= lf.ptr + lf.bytes) {
            Temp &t = g_rec.plan.temps[it->second.second];
            if (t.up_leaf >= 0 || n > t.bytes) {
                g_rec.bad = true;
                g_rec.why = "second upload into one temporary";
                return;
            }
            t.up_leaf = (int)k;
            t.up_off = (size_t)(s - lf.ptr);
            t.up_bytes = n;
            g_rec.rec_h2d++;
            return;
        }
    }
    g_rec.bad = true;
    g_rec.why = "upload from host memory that is not a leaf of the graph";
}
void note_foreign(const char *why) {
    if (g_rec.active) {
        g_rec.bad = true;
        g_rec.why = why;
    }
}

bool try_replay(const graph::GraphNode *root, Tensor *out) {
    g_pending.valid = false;
    if (!enabled())
        return false;
    if (g_rec.active) { // a materialisation inside a recorded lowering: the outer recording cannot express it
        note_foreign("nested materialisation");
        return false;
    }
    std::vector<Leaf> leaves;
    uint64_t lh = 0;
    if (!collect(root, &leaves, &lh))
        return false;
    const uint64_t key = mix(graph::compute_signature(root).hash, lh);
    std::shared_ptr<Plan> pl;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) {
            pl = it->second;
            pl->tick = ++g_tick;
        }
    }
    g_pending.valid = true;
    g_pending.root = root;
    g_pending.key = key;
    g_pending.known_uncacheable = pl && !pl->cacheable;
    if (!pl || !pl->cacheable || pl->n_leaves != leaves.size()) {
        if (pl && !pl->cacheable) {
            std::lock_guard<std::mutex> lk(g_mu);
            g_stats.uncacheable++;
        }
        g_pending.leaves = std::move(leaves);
        return false;
    }
    // ---- replay --------------------------------------------------------------------------------------------------
    const Plan &P = *pl;
    std::vector<void *> tptr(P.temps.size(), nullptr);
    const int result_temp = P.result.kind == K_TEMP ? P.result.idx : -1;
    auto release_all = [&](bool keep_result) {
        for (size_t j = 0; j < tptr.size(); ++j)
            if (tptr[j] && !(keep_result && (int)j == result_temp)) {
                axb_free(tptr[j], nullptr);
                tptr[j] = nullptr;
            }
    };
    auto ensure = [&](int j) {
        if (tptr[j])
            return;
        const Temp &t = P.temps[j];
        check(axb_alloc(t.bytes, nullptr, &tptr[j]), "axb_alloc (plan replay)");
        if (t.up_leaf >= 0) // host scalar / small constant: the CURRENT leaf's bytes
            check(axb_memcpy_h2d(tptr[j], leaves[t.up_leaf].ptr + t.up_off, t.up_bytes, nullptr), "axb_memcpy_h2d (plan replay)");
    };
    try {
        axb_tensor T[kMaxT];
        for (size_t i = 0; i < P.calls.size(); ++i) {
            const Call &c = P.calls[i];
            for (int k = 0; k < kMaxT; ++k) {
                const TRef &r = c.t[k];
                if (r.kind == K_NONE)
                    continue;
                T[k] = r.d;
                if (r.kind == K_LEAF) {
                    T[k].base = const_cast<char *>(leaves[r.idx].ptr);
                } else {
                    ensure(r.idx);
                    T[k].base = tptr[r.idx];
                }
            }
            check(invoke(c, T), "plan replay");
            for (int k = 0; k < kMaxT; ++k) { // stream-ordered allocator: a temporary goes back right after its last use
                const TRef &r = c.t[k];
                if (r.kind == K_TEMP && P.temps[r.idx].last == (int)i && r.idx != result_temp && tptr[r.idx]) {
                    axb_free(tptr[r.idx], nullptr);
                    tptr[r.idx] = nullptr;
                }
            }
        }
        std::shared_ptr<Storage> st;
        if (P.result.kind == K_LEAF) {
            st = leaves[P.result.idx].st;
        } else {
            ensure(result_temp);
            st = std::make_shared<CudaStorage>(tptr[result_temp], P.temps[result_temp].bytes, CudaStorage::Adopt{});
            tptr[result_temp] = nullptr; // owned by the storage now
        }
        *out = Tensor(st, P.shape, P.strides, P.dtype, (size_t)P.result.d.byte_offset);
    } catch (...) {
        release_all(false);
        throw;
    }
    release_all(false);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_stats.hits++;
    }
    g_pending.valid = false;
    return true;
}

void begin_record(const graph::GraphNode *root) {
    if (!enabled() || g_rec.active || !g_pending.valid || g_pending.root != root)
        return;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_pending.known_uncacheable)
            g_stats.misses++;
    }
    if (g_pending.known_uncacheable) {
        g_pending.valid = false;
        return;
    }
    g_rec = Rec{};
    g_rec.active = true;
    g_rec.key = g_pending.key;
    g_rec.leaves = std::move(g_pending.leaves);
    g_pending.valid = false;
    for (size_t k = 0; k < g_rec.leaves.size(); ++k)
        if (g_rec.leaves[k].gpu && g_rec.leaves[k].ptr)
            g_rec.slot.emplace(static_cast<const void *>(g_rec.leaves[k].ptr), std::make_pair((int8_t)K_LEAF, (int32_t)k));
    g_rec.plan.n_leaves = g_rec.leaves.size();
    g_rec.launches0 = axb_launch_count();
    axb_get_transfer_stats(&g_rec.x0);
}

void end_record(const Tensor *result, bool ok) {
    if (!g_rec.active)
        return;
    Rec rec = std::move(g_rec);
    g_rec = Rec{};
    bool good = ok && !rec.bad && result && result->storage();
    if (good) { // nothing may have escaped the recording
        axb_transfer_stats x1{};
        axb_get_transfer_stats(&x1);
        good = axb_launch_count() - rec.launches0 == rec.rec_launches && x1.h2d_calls - rec.x0.h2d_calls == rec.rec_h2d &&
               x1.d2h_calls == rec.x0.d2h_calls && x1.d2d_calls == rec.x0.d2d_calls;
    }
    auto plan = std::make_shared<Plan>(std::move(rec.plan));
    if (good) {
        auto *cs = dynamic_cast<CudaStorage *>(result->storage().get());
        auto it = cs ? rec.slot.find(cs->device_ptr()) : rec.slot.end();
        if (it == rec.slot.end()) {
            good = false;
        } else {
            plan->result.kind = it->second.first;
            plan->result.idx = it->second.second;
            plan->result.d.byte_offset = (int64_t)result->offset();
            plan->shape = result->shape();
            plan->strides = result->strides();
            plan->dtype = result->dtype();
            // the address a leaf and a temporary share is ambiguous only if the result IS that address; leaves win in
            // `slot` unless a temporary re-registered it, which cannot happen while the leaf's storage is alive
        }
    }
    if (good) {
        for (size_t i = 0; i < plan->calls.size(); ++i)
            for (int k = 0; k < kMaxT; ++k)
                if (plan->calls[i].t[k].kind == K_TEMP)
                    plan->temps[plan->calls[i].t[k].idx].last = (int)i;
        // a temporary that is only uploaded into but never read is dead; one that is read without ever being written by a
        // recorded call or upload would be garbage — the recorder only sees allocations made through the backend, all of
        // which are followed by the call that fills them, so this cannot happen for a consistent recording.
    } else {
        plan->calls.clear();
        plan->temps.clear();
    }
    plan->cacheable = good;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!good)
        g_stats.uncacheable++;
    if (g_cache.size() >= kMaxPlans) { // evict the least recently used entry
        auto victim = g_cache.begin();
        for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
            if (it->second->tick < victim->second->tick)
                victim = it;
        g_cache.erase(victim);
        g_stats.evictions++;
    }
    plan->tick = ++g_tick;
    g_cache[rec.key] = std::move(plan);
}

// ---- wrappers ----------------------------------------------------------------------------------------------------------
#define AXB_PLAN_CALL(expr)                                                                                                     \
    Scope sc;                                                                                                                   \
    const axb_status st_ = (expr);                                                                                              \
    if (st_ != AXB_OK || !sc.on)                                                                                                \
        return st_;

axb_status elementwise(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_elementwise(prog, n_instr, inputs, n_inputs, out, nullptr))
    Call *c = sc.open(F_ELEMENTWISE);
    c->prog.assign(prog, prog + n_instr);
    c->iv[0] = n_inputs;
    for (int i = 0; i < n_inputs; ++i)
        c->t[i] = resolve(&inputs[i]);
    c->t[n_inputs] = resolve(out);
    return st_;
}
axb_status elementwise_reduce(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t reduce_op,
                              const axb_tensor *out) {
    AXB_PLAN_CALL(axb_elementwise_reduce(prog, n_instr, inputs, n_inputs, reduce_op, out, nullptr))
    Call *c = sc.open(F_ELEMENTWISE_REDUCE);
    c->prog.assign(prog, prog + n_instr);
    c->iv[0] = n_inputs;
    c->iv[1] = reduce_op;
    for (int i = 0; i < n_inputs; ++i)
        c->t[i] = resolve(&inputs[i]);
    c->t[n_inputs] = resolve(out);
    return st_;
}
axb_status elementwise_reduce_axis(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t reduce_op,
                                   int axis, int keep_dims, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_elementwise_reduce_axis(prog, n_instr, inputs, n_inputs, reduce_op, axis, keep_dims, out, nullptr))
    Call *c = sc.open(F_ELEMENTWISE_REDUCE_AXIS);
    c->prog.assign(prog, prog + n_instr);
    c->iv[0] = n_inputs;
    c->iv[1] = reduce_op;
    c->iv[2] = axis;
    c->iv[3] = keep_dims;
    for (int i = 0; i < n_inputs; ++i)
        c->t[i] = resolve(&inputs[i]);
    c->t[n_inputs] = resolve(out);
    return st_;
}
axb_status elementwise_softmax(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int is_log, int32_t epilogue_op,
                               const axb_tensor *out) {
    AXB_PLAN_CALL(axb_elementwise_softmax(prog, n_instr, inputs, n_inputs, is_log, epilogue_op, out, nullptr))
    Call *c = sc.open(F_ELEMENTWISE_SOFTMAX);
    c->prog.assign(prog, prog + n_instr);
    c->iv[0] = n_inputs;
    c->iv[1] = is_log;
    c->iv[2] = epilogue_op;
    for (int i = 0; i < n_inputs; ++i)
        c->t[i] = resolve(&inputs[i]);
    c->t[n_inputs] = resolve(out);
    return st_;
}
axb_status elementwise_norm(const axb_instr *prog, int n_instr, const axb_tensor *inputs, int n_inputs, int32_t kind, const axb_tensor *w,
                            const axb_tensor *b, float eps, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_elementwise_norm(prog, n_instr, inputs, n_inputs, kind, w, b, eps, out, nullptr))
    if (n_inputs + 3 > kMaxT) { // cannot be expressed in one recorded call: this signature keeps the normal path
        note_foreign("elementwise_norm with more operands than a recorded call holds");
        return st_;
    }
    Call *c = sc.open(F_ELEMENTWISE_NORM);
    c->prog.assign(prog, prog + n_instr);
    c->iv[0] = n_inputs;
    c->iv[1] = kind;
    c->dv[0] = eps;
    for (int i = 0; i < n_inputs; ++i)
        c->t[i] = resolve(&inputs[i]);
    c->t[n_inputs] = resolve(out);
    c->t[n_inputs + 1] = resolve(w);
    c->t[n_inputs + 2] = resolve(b);
    return st_;
}
axb_status matmul(const axb_tensor *a, const axb_tensor *b, int ta, int tb, const axb_matmul_epilogue *epi, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_matmul(a, b, ta, tb, epi, out, nullptr))
    Call *c = sc.open(F_MATMUL);
    c->iv[0] = ta; c->iv[1] = tb;
    c->iv[2] = epi ? epi->activation : 0;
    c->iv[3] = epi ? epi->flags : 0;
    c->iv[4] = epi ? 1 : 0;
    c->t[0] = resolve(a);
    c->t[1] = resolve(b);
    c->t[2] = resolve(epi ? epi->bias : nullptr);
    c->t[3] = resolve(out);
    return st_;
}
axb_status softmax(const axb_tensor *in, int axis, int is_log, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_softmax(in, axis, is_log, out, nullptr))
    Call *c = sc.open(F_SOFTMAX);
    c->iv[0] = axis; c->iv[1] = is_log;
    c->t[0] = resolve(in);
    c->t[1] = resolve(out);
    return st_;
}
axb_status layer_norm(const axb_tensor *in, const axb_tensor *w, const axb_tensor *b, int axis, float eps, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_layer_norm(in, w, b, axis, eps, out, nullptr))
    Call *c = sc.open(F_LAYER_NORM);
    c->iv[0] = axis; c->dv[0] = eps;
    c->t[0] = resolve(in); c->t[1] = resolve(w); c->t[2] = resolve(b); c->t[3] = resolve(out);
    return st_;
}
axb_status rms_norm(const axb_tensor *in, const axb_tensor *w, int axis, float eps, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_rms_norm(in, w, axis, eps, out, nullptr))
    Call *c = sc.open(F_RMS_NORM);
    c->iv[0] = axis; c->dv[0] = eps;
    c->t[0] = resolve(in); c->t[1] = resolve(w); c->t[2] = resolve(out);
    return st_;
}
axb_status reduce(int32_t op, const axb_tensor *in, const int32_t *axes, int n_axes, int keep_dims, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_reduce(op, in, axes, n_axes, keep_dims, out, nullptr))
    Call *c = sc.open(F_REDUCE);
    c->iv[0] = op; c->iv[1] = n_axes; c->iv[2] = keep_dims;
    for (int i = 0; i < n_axes && i < AXB_MAX_DIMS; ++i)
        c->iv[3 + i] = axes[i];
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status argreduce(int32_t op, const axb_tensor *in, int axis, int keep_dims, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_argreduce(op, in, axis, keep_dims, out, nullptr))
    Call *c = sc.open(F_ARGREDUCE);
    c->iv[0] = op; c->iv[1] = axis; c->iv[2] = keep_dims;
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status cast(const axb_tensor *in, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_cast(in, out, nullptr))
    Call *c = sc.open(F_CAST);
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status gather_strided(const axb_tensor *in, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_gather_strided(in, out, nullptr))
    Call *c = sc.open(F_GATHER_STRIDED);
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status im2col_1d(const axb_tensor *in, int k, int stride, int padding, int dilation, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_im2col_1d(in, k, stride, padding, dilation, out, nullptr))
    Call *c = sc.open(F_IM2COL);
    c->iv[0] = k; c->iv[1] = stride; c->iv[2] = padding; c->iv[3] = dilation;
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status binary(int32_t op, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_binary(op, a, b, out, nullptr))
    Call *c = sc.open(F_BINARY);
    c->iv[0] = op;
    c->t[0] = resolve(a); c->t[1] = resolve(b); c->t[2] = resolve(out);
    return st_;
}
axb_status unary(int32_t op, const axb_tensor *in, const axb_tensor *out, double param) {
    AXB_PLAN_CALL(axb_unary(op, in, out, param, nullptr))
    Call *c = sc.open(F_UNARY);
    c->iv[0] = op; c->dv[0] = param;
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status where(const axb_tensor *cnd, const axb_tensor *a, const axb_tensor *b, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_where(cnd, a, b, out, nullptr))
    Call *c = sc.open(F_WHERE);
    c->t[0] = resolve(cnd); c->t[1] = resolve(a); c->t[2] = resolve(b); c->t[3] = resolve(out);
    return st_;
}
axb_status masked_fill(const axb_tensor *in, const axb_tensor *mask, double value, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_masked_fill(in, mask, value, out, nullptr))
    Call *c = sc.open(F_MASKED_FILL);
    c->dv[0] = value;
    c->t[0] = resolve(in); c->t[1] = resolve(mask); c->t[2] = resolve(out);
    return st_;
}
axb_status pad_constant(const axb_tensor *in, const int64_t *lo, const int64_t *hi, double value, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_pad_constant(in, lo, hi, value, out, nullptr))
    Call *c = sc.open(F_PAD);
    for (int i = 0; i < in->ndim && i < 8; ++i) {
        c->iv[i] = lo[i];
        c->iv[8 + i] = hi[i];
    }
    c->dv[0] = value;
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status conv(const axb_tensor *in, const axb_tensor *w, const axb_tensor *b, int nsp, const int32_t *stride, const int32_t *padding,
                const int32_t *dilation, int groups, int activation, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_conv(in, w, b, nsp, stride, padding, dilation, groups, activation, out, nullptr))
    Call *c = sc.open(F_CONV2D);
    c->iv[0] = stride[0]; c->iv[1] = nsp > 1 ? stride[1] : 1; c->iv[2] = padding[0]; c->iv[3] = nsp > 1 ? padding[1] : 0;
    c->iv[4] = dilation[0]; c->iv[5] = nsp > 1 ? dilation[1] : 1; c->iv[6] = groups; c->iv[7] = nsp; c->iv[8] = activation;
    c->t[0] = resolve(in); c->t[1] = resolve(w); c->t[2] = resolve(b); c->t[3] = resolve(out);
    return st_;
}
axb_status pool(const axb_tensor *in, int is_avg, int nsp, const int32_t *kernel, const int32_t *stride, const int32_t *padding,
                int count_include_pad, const axb_tensor *out) {
    AXB_PLAN_CALL(axb_pool(in, is_avg, nsp, kernel, stride, padding, count_include_pad, out, nullptr))
    Call *c = sc.open(F_POOL);
    c->iv[0] = is_avg; c->iv[1] = nsp; c->iv[2] = count_include_pad;
    for (int d = 0; d < nsp && d < 3; ++d) {
        c->iv[3 + d] = kernel[d];
        c->iv[6 + d] = stride[d];
        c->iv[9 + d] = padding[d];
    }
    c->t[0] = resolve(in); c->t[1] = resolve(out);
    return st_;
}
axb_status sort(const axb_tensor *in, int axis, int descending, const axb_tensor *values, const axb_tensor *indices) {
    AXB_PLAN_CALL(axb_sort(in, axis, descending, values, indices, nullptr))
    Call *c = sc.open(F_SORT);
    c->iv[0] = axis; c->iv[1] = descending;
    c->t[0] = resolve(in); c->t[1] = resolve(values); c->t[2] = resolve(indices);
    return st_;
}
axb_status fill(const axb_tensor *out, double value) {
    AXB_PLAN_CALL(axb_fill(out, value, nullptr))
    Call *c = sc.open(F_FILL);
    c->dv[0] = value;
    c->t[0] = resolve(out);
    return st_;
}

} // namespace plan
} // namespace cuda
} // namespace backends
} // namespace axiom
