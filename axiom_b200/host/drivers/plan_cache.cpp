// plan_cache — hit / miss behaviour of the graph-level plan cache of the CUDA lowering (host/cuda_plan.cpp),
// written the way the reference tests its own cache (tests/test_graph_compiler.cpp:79-125 CacheHit / CacheMissOnShapeChange):
// every case materialises the same Tensor-API expression on Device::GPU several times and reads the counters.
//   * same expression, same layouts, NEW data and NEW scalar values   -> hit, result still equal to Device::CPU
//   * same expression on a strided / transposed / offset view         -> its own plan (the reference's signature ignores
//                                                                        leaf strides, graph_signature.cpp:129-139)
//   * shape change, dtype change, aliasing change (f(x, x) vs f(x, y)) -> miss
//   * a replay launches exactly the kernels the recording launched and moves no device->host bytes
// One JSON object on stdout; exit code 0 = all ok.
#include <axiom/axiom.hpp>

#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "axb.h"
#include "cuda_plan.hpp"

using namespace axiom;
namespace plan = axiom::backends::cuda::plan;

struct Case {
    std::string name;
    bool ok;
    std::string why;
};
static std::vector<Case> g_cases;
static void report(const std::string &name, bool ok, const std::string &why = "") { g_cases.push_back({name, ok, ok ? "" : why}); }

static bool same_bytes(const Tensor &gpu, const Tensor &cpu) {
    Tensor a = gpu.cpu().ascontiguousarray(), b = cpu.ascontiguousarray();
    return a.shape() == b.shape() && a.dtype() == b.dtype() && std::memcmp(a.data(), b.data(), b.nbytes()) == 0;
}

struct Run {
    uint64_t hits, misses, launches, d2h;
    Tensor out;
};
static std::string g_trace;
static Run run(const std::function<Tensor()> &f) {
    const plan::Stats s0 = plan::stats();
    axb_transfer_stats t0{}, t1{};
    axb_get_transfer_stats(&t0);
    const uint64_t l0 = axb_launch_count();
    Tensor r = f();
    r.sync();
    axb_get_transfer_stats(&t1);
    const plan::Stats s1 = plan::stats();
    char buf[160];
    std::snprintf(buf, sizeof buf, "[h%llu m%llu u%llu l%llu]", (unsigned long long)(s1.hits - s0.hits),
                  (unsigned long long)(s1.misses - s0.misses), (unsigned long long)(s1.uncacheable - s0.uncacheable),
                  (unsigned long long)(axb_launch_count() - l0));
    g_trace += buf;
    return {s1.hits - s0.hits, s1.misses - s0.misses, axb_launch_count() - l0, t1.d2h_bytes - t0.d2h_bytes, r};
}

int main() {
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(4242);
    if (!plan::enabled()) {
        std::printf("{\"all_ok\": false, \"failed\": [{\"name\": \"enabled\", \"why\": \"plan cache switched off\"}]}\n");
        return 1;
    }
    plan::clear();

    // ---- 1. CacheHit: elementwise chain with scalar operands, new data + new scalar each time --------------------------
    auto chain = [](const Tensor &x, const Tensor &y, float s) { return ((x.relu() + s) * y - x).abs(); };
    {
        Tensor x1 = Tensor::randn({257, 129}), y1 = Tensor::randn({1, 129}), x2 = Tensor::randn({257, 129}) * 3.0f,
               y2 = Tensor::randn({1, 129});
        Tensor x1g = x1.gpu(), y1g = y1.gpu(), x2g = x2.gpu(), y2g = y2.gpu();
        Run a = run([&] { return chain(x1g, y1g, 1.0f); });
        Run b = run([&] { return chain(x2g, y2g, 2.5f); });
        Run c = run([&] { return chain(x1g, y1g, -0.75f); });
        report("chain_first_is_miss", a.misses == 1 && a.hits == 0, "expected one miss");
        report("chain_second_is_hit", b.hits == 1 && b.misses == 0, "expected one hit");
        report("chain_third_is_hit", c.hits == 1 && c.misses == 0, "expected one hit");
        report("chain_replay_same_launches", a.launches == b.launches && a.launches == 1, "launch counts differ");
        report("chain_no_d2h", a.d2h == 0 && b.d2h == 0, "device->host bytes inside a GPU call");
        Tensor e1 = chain(x1, y1, 1.0f), e2 = chain(x2, y2, 2.5f), e3 = chain(x1, y1, -0.75f);
        report("chain_recorded_bit_exact", same_bytes(a.out, e1), "recorded run differs from the CPU backend");
        report("chain_replayed_bit_exact_new_data", same_bytes(b.out, e2), "replay used stale data or a stale scalar");
        report("chain_replayed_bit_exact_new_scalar", same_bytes(c.out, e3), "replay used a stale scalar");

        // ---- 2. layout classes: transposed, strided-slice and offset views get their own plans ------------------------
        Tensor xt = Tensor::randn({129, 257});
        Tensor xtg = xt.gpu();
        Run t1 = run([&] { return chain(xtg.T(), y1g, 1.0f); });
        Run t2 = run([&] { return chain(xtg.T(), y2g, 0.5f); });
        report("transposed_view_is_new_plan", t1.misses == 1 && t1.hits == 0, "a transposed leaf reused the contiguous plan");
        report("transposed_view_then_hit", t2.hits == 1, "expected a hit");
        report("transposed_bit_exact", same_bytes(t1.out, chain(xt.T(), y1, 1.0f)) && same_bytes(t2.out, chain(xt.T(), y2, 0.5f)),
               "transposed view differs from the CPU backend");
        Tensor big = Tensor::randn({514, 129});
        Tensor bigg = big.gpu();
        auto ev = [](const Tensor &t) { return t.slice({Slice(0, 514, 2), Slice()}); };
        auto lo = [](const Tensor &t) { return t.slice({Slice(257, 514), Slice()}); };
        Run s1 = run([&] { return chain(ev(bigg), y1g, 1.0f); });
        Run s2 = run([&] { return chain(lo(bigg), y1g, 1.0f); });
        Run s3 = run([&] { return chain(ev(bigg), y2g, 3.0f); });
        report("strided_view_is_new_plan", s1.misses == 1, "a strided leaf reused another plan");
        report("strided_and_offset_bit_exact",
               same_bytes(s1.out, chain(ev(big), y1, 1.0f)) && same_bytes(s2.out, chain(lo(big), y1, 1.0f)) &&
                   same_bytes(s3.out, chain(ev(big), y2, 3.0f)),
               "view result differs from the CPU backend");
        report("strided_view_then_hit", s3.hits == 1, "expected a hit");

        // ---- 3. CacheMissOnShapeChange / dtype change --------------------------------------------------------------------
        Tensor x3 = Tensor::randn({64, 129});
        Tensor x3g = x3.gpu();
        Run sh = run([&] { return chain(x3g, y1g, 1.0f); });
        report("shape_change_is_miss", sh.misses == 1 && sh.hits == 0, "a different shape hit the cache");
        report("shape_change_bit_exact", same_bytes(sh.out, chain(x3, y1, 1.0f)), "differs from the CPU backend");
        Tensor xh = x1.astype(DType::BFloat16), yh = y1.astype(DType::BFloat16);
        xh.sync();
        yh.sync();
        Tensor xhg = xh.gpu(), yhg = yh.gpu();
        Run dt = run([&] { return chain(xhg, yhg, 1.0f); });
        report("dtype_change_is_miss", dt.misses == 1 && dt.hits == 0, "a different dtype hit the cache");
        report("dtype_change_bit_exact", same_bytes(dt.out, chain(xh, yh, 1.0f)), "bf16 chain differs from the CPU backend");
    }

    // ---- 4. aliasing is part of the key: f(x, x) and f(x, y) are different plans ------------------------------------------
    {
        auto f = [](const Tensor &p, const Tensor &q) { return (p * q + p).tanh(); };
        Tensor p = Tensor::randn({96, 80}), q = Tensor::randn({96, 80});
        Tensor pg = p.gpu(), qg = q.gpu();
        Run a = run([&] { return f(pg, pg); });
        Run b = run([&] { return f(pg, qg); });
        Run c = run([&] { return f(qg, pg); });
        Run d = run([&] { return f(qg, qg); });
        report("alias_plans_distinct", a.misses == 1 && b.misses == 1, "f(x,x) and f(x,y) shared a plan");
        report("alias_hits", c.hits == 1 && d.hits == 1, "expected hits for the same alias structure");
        report("alias_bit_exact", same_bytes(a.out, f(p, p)) && same_bytes(b.out, f(p, q)) && same_bytes(c.out, f(q, p)) &&
                                      same_bytes(d.out, f(q, q)),
               "aliasing replay differs from the CPU backend");
    }

    // ---- 5. a multi-kernel plan with temporaries: config 1 (matmul -> softmax -> sum) and a norm / reduce mix -----------
    {
        auto c1 = [](const Tensor &t) { return t.matmul(t.T()).softmax(-1).sum({1}); };
        Tensor u = Tensor::randn({192, 160}), v = Tensor::randn({192, 160}) * 0.5f;
        Tensor ug = u.gpu(), vg = v.gpu();
        Run a = run([&] { return c1(ug); });
        Run b = run([&] { return c1(vg); });
        report("config1_hit", a.misses == 1 && b.hits == 1 && b.misses == 0, "config-1 chain: expected miss then hit");
        report("config1_same_launches", a.launches == b.launches && a.launches >= 2, "replay launched a different kernel count");
        report("config1_parity", c1(u).allclose(a.out.cpu(), 1e-5, 1e-6) && c1(v).allclose(b.out.cpu(), 1e-5, 1e-6),
               "config-1 chain outside 1e-5 of the CPU backend");
        // the replay of a deterministic kernel sequence is bit-identical to an unrecorded run of the same data
        plan::clear();
        Run b2 = run([&] { return c1(vg); });
        report("config1_replay_equals_fresh_lowering", b2.misses == 1 && same_bytes(b.out, b2.out.cpu()),
               "replayed result differs from a fresh lowering");

        auto mix = [](const Tensor &t, const Tensor &w, const Tensor &bb) {
            return (ops::layer_norm(t, w, bb, -1, 1e-5f).gelu() * 2.0f).max(1);
        };
        Tensor w = Tensor::randn({160}), bb = Tensor::randn({160});
        Tensor wg = w.gpu(), bg = bb.gpu();
        Run m1 = run([&] { return mix(ug, wg, bg); });
        Run m2 = run([&] { return mix(vg, wg, bg); });
        report("norm_mix_hit", m2.hits >= 1 && m2.misses == 0, "layer_norm chain: expected a hit on the second call");
        report("norm_mix_parity", mix(u, w, bb).allclose(m1.out.cpu(), 1e-5, 1e-5) && mix(v, w, bb).allclose(m2.out.cpu(), 1e-5, 1e-5),
               "layer_norm chain outside 1e-5 of the CPU backend");
        report("norm_mix_no_d2h", m1.d2h == 0 && m2.d2h == 0, "device->host bytes inside a GPU call");
    }

    // ---- 6. where-chains are lowered by fused_where (cuda_graph.cpp) outside materialize_gpu_graph: one launch, served by
    //         the per-launch plan cache of the fused-chain family (axb_get_plan_cache_stats) -------------------------------
    {
        auto wc = [](const Tensor &x, const Tensor &y, float s, int z) { return ((x.relu() + s) * y).where(x > 0, z); };
        Tensor x = Tensor::randn({130, 66}), y = Tensor::randn({130, 1});
        Tensor xg = x.gpu(), yg = y.gpu();
        axb_plan_cache_stats p0{}, p1{}, p2{};
        axb_get_plan_cache_stats(&p0);
        Run a = run([&] { return wc(xg, yg, 1.0f, 0); });
        axb_get_plan_cache_stats(&p1);
        Run b = run([&] { return wc(xg, yg, -3.0f, 7); });
        axb_get_plan_cache_stats(&p2);
        report("where_chain_one_launch", a.launches == 1 && b.launches == 1, "where-chain is not one kernel");
        report("where_chain_launch_plan_miss_then_hit", p1.misses - p0.misses == 1 && p2.hits - p1.hits == 1 && p2.misses == p1.misses,
               "launch-plan cache: expected a miss then a hit");
        report("where_chain_bit_exact", same_bytes(a.out, wc(x, y, 1.0f, 0)) && same_bytes(b.out, wc(x, y, -3.0f, 7)),
               "where-chain differs from the CPU backend");
    }

    // ---- 7. counters and clear() ------------------------------------------------------------------------------------------
    {
        const plan::Stats s = plan::stats();
        report("entries_counted", s.entries >= 2 && s.entries <= 512, "entries out of range");
        plan::clear();
        const plan::Stats z = plan::stats();
        report("clear_resets", z.entries == 0 && z.hits == 0 && z.misses == 0, "clear() left state behind");
    }

    bool all = true;
    std::printf("{\"cases\": %zu, \"failed\": [", g_cases.size());
    bool first = true;
    for (auto &k : g_cases)
        if (!k.ok) {
            all = false;
            std::printf("%s{\"name\": \"%s\", \"why\": \"%s\"}", first ? "" : ", ", k.name.c_str(), k.why.c_str());
            first = false;
        }
    std::printf("], \"trace\": \"%s\", \"all_ok\": %s}\n", g_trace.c_str(), all ? "true" : "false");
    return all ? 0 : 1;
}
