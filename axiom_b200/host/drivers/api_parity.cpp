// api_parity — small GPU-vs-CPU parity cases through the reference's Tensor API that the reference's own tests do not hold
// (leaky_relu on the lazy GPU path silently became relu). Every case runs the SAME expression on Device::CPU
// and Device::GPU of the same library and compares the results; fp32 elementwise is bit-exact, half types exact as well
// (both sides compute in float and round once, bfloat16.hpp:55-66).
// One JSON object on stdout; exit code 0 = all ok.
#include <axiom/axiom.hpp>

#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "axb.h"

using namespace axiom;

struct Case {
    std::string name;
    bool ok;
    std::string why;
};

int main() {
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(99);
    std::vector<Case> cases;
    auto check = [&](const std::string &name, const Tensor &x_cpu, const std::function<Tensor(const Tensor &)> &f) {
        Case c{name, false, ""};
        try {
            Tensor ref = f(x_cpu);
            ref.sync();
            Tensor xg = x_cpu.gpu();
            axb_transfer_stats t0{}, t1{};
            axb_get_transfer_stats(&t0);
            const uint64_t l0 = axb_launch_count();
            Tensor got = f(xg);
            got.sync();
            axb_get_transfer_stats(&t1);
            if (got.device() != Device::GPU)
                c.why = "result is not on the GPU";
            else if (axb_launch_count() == l0)
                c.why = "no kernel launched";
            else if (t1.d2h_bytes != t0.d2h_bytes)
                c.why = "device->host traffic inside the GPU call (CPU fallback)";
            else {
                Tensor back = got.cpu().ascontiguousarray();
                Tensor r = ref.ascontiguousarray();
                if (back.shape() != r.shape() || back.dtype() != r.dtype())
                    c.why = "shape / dtype differ";
                else if (std::memcmp(back.data(), r.data(), r.nbytes()) != 0)
                    c.why = "values differ from the CPU backend";
                else
                    c.ok = true;
            }
        } catch (const std::exception &e) {
            c.why = e.what();
        }
        cases.push_back(c);
    };

    Tensor x = Tensor::randn({37, 129}) * 3.0f;
    x.sync();
    for (DType dt : {DType::Float32, DType::BFloat16, DType::Float16}) {
        const std::string sfx = dt == DType::Float32 ? "f32" : dt == DType::BFloat16 ? "bf16" : "f16";
        Tensor xd = x.astype(dt);
        xd.sync();
        // default slope through the LAZY unary node (operations.cpp:1391-1394: NoParams)
        check("leaky_relu_default_" + sfx, xd, [](const Tensor &t) { return ops::leaky_relu(t); });
        check("leaky_relu_method_" + sfx, xd, [](const Tensor &t) { return t.leaky_relu(); });
        // custom slope: relu(x) - slope * relu(-x) (operations.cpp:1396-1400)
        check("leaky_relu_0p2_" + sfx, xd, [](const Tensor &t) { return ops::leaky_relu(t, 0.2f); });
        // inside a longer lazy chain, and under a full reduction (the chain + reduction fusion path)
        check("leaky_relu_chain_" + sfx, xd, [](const Tensor &t) { return (ops::leaky_relu(t) + 1.0f) * t; });
    }
    // elementwise producer + reduction over the last axis: ONE kernel (the reference fuses producers into reductions,
    // graph_compiler.cpp:376-478); sums within the 1e-5 contract, max / min exact
    {
        Tensor p = Tensor::randn({257, 1500}), q = Tensor::randn({1, 1500}), sm = Tensor::randn({600, 200});
        p.sync(); q.sync(); sm.sync();
        auto fused = [&](const std::string &name, const Tensor &a_cpu, const std::function<Tensor(const Tensor &)> &f, bool exact) {
            Case c{name, false, ""};
            try {
                Tensor ref = f(a_cpu);
                ref.sync();
                Tensor ag = a_cpu.gpu();
                const uint64_t l0 = axb_launch_count();
                Tensor got = f(ag);
                got.sync();
                const uint64_t launches = axb_launch_count() - l0;
                Tensor back = got.cpu();
                if (got.device() != Device::GPU)
                    c.why = "result is not on the GPU";
                else if (launches != 1)
                    c.why = "expected one fused kernel, launched " + std::to_string(launches);
                else if (back.shape() != ref.shape())
                    c.why = "shape differs";
                else if (exact ? std::memcmp(back.ascontiguousarray().data(), ref.ascontiguousarray().data(), ref.nbytes()) != 0
                               : !ref.allclose(back, 1e-4, 1e-4))
                    c.why = "values differ from the CPU backend";
                else
                    c.ok = true;
            } catch (const std::exception &e) {
                c.why = e.what();
            }
            cases.push_back(c);
        };
        Tensor qg = q.gpu();
        auto Q = [&](const Tensor &t) { return t.device() == Device::GPU ? qg : q; };
        fused("chain_sum_last_axis", p, [&](const Tensor &t) { return (t * Q(t) + 1.0f).sum({1}); }, false);
        fused("chain_mean_last_axis_keepdims", p, [&](const Tensor &t) { return t.square().mean({1}, true); }, false);
        fused("chain_max_last_axis", p, [&](const Tensor &t) { return (t * Q(t)).abs().max(1); }, true);
        fused("chain_min_short_rows", sm, [&](const Tensor &t) { return (t.relu() - 0.5f).min(1); }, true);
        // softmax with an elementwise prologue, and with a row-reduction epilogue
        fused("chain_softmax", p, [&](const Tensor &t) { return (t * 0.5f + Q(t)).softmax(-1); }, false);
        fused("chain_log_softmax", sm, [&](const Tensor &t) { return (t * t - 1.0f).log_softmax(-1); }, false);
        fused("softmax_then_row_max", sm, [&](const Tensor &t) { return (t * 2.0f).softmax(-1).max(1); }, false);
        // layer_norm / rms_norm with an elementwise prologue (post-LN: layer_norm(x + sublayer(x)))
        Tensor lw = Tensor::randn({1500}) * 0.1f + 1.0f, lb = Tensor::randn({1500}) * 0.1f;
        lw.sync(); lb.sync();
        Tensor lwg = lw.gpu(), lbg = lb.gpu();
        auto W = [&](const Tensor &t) { return t.device() == Device::GPU ? lwg : lw; };
        auto Bv = [&](const Tensor &t) { return t.device() == Device::GPU ? lbg : lb; };
        fused("chain_layer_norm", p, [&](const Tensor &t) { return ops::layer_norm(t + Q(t), W(t), Bv(t), -1, 1e-5f); }, false);
        fused("chain_rms_norm", p, [&](const Tensor &t) { return ops::rms_norm(t * 0.5f + 1.0f, W(t), -1, 1e-6f); }, false);
    }
    // scalar-operand forms of where (the patched Tensor::where(cond, scalar) overloads, tensor.cpp:1340-1355)
    Tensor y = Tensor::randn({5, 1, 33}), c = Tensor::randn({5, 7, 33});
    y.sync();
    c.sync();
    auto with_c = [&](const Tensor &t) { return t.device() == Device::GPU ? c.gpu() : c; };
    check("where_scalar_int", y, [&](const Tensor &t) { return t.where(with_c(t) > 0, 0); });
    check("where_scalar_float", y, [&](const Tensor &t) { return t.where(with_c(t) > 0.5f, -2.5f); });
    check("where_scalar_double", y, [&](const Tensor &t) { return t.where(with_c(t) <= 0.0, 7.0); });

    bool all = true;
    std::printf("{\"cases\": %zu, \"failed\": [", cases.size());
    bool first = true;
    for (auto &k : cases)
        if (!k.ok) {
            all = false;
            std::printf("%s{\"name\": \"%s\", \"why\": \"%s\"}", first ? "" : ", ", k.name.c_str(), k.why.c_str());
            first = false;
        }
    std::printf("], \"all_ok\": %s}\n", all ? "true" : "false");
    return all ? 0 : 1;
}
