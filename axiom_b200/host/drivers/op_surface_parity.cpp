// op_surface_parity — pooling, sort / argsort / topk and masked_select through the reference's Tensor
// API on Device::GPU against the SAME call on Device::CPU of the same library. The reference has no device branch for these
// ops (a GPU tensor makes a host round trip, src/tensor/operations.cpp:2367-2880, :4199-4257); the patched build routes them
// to csrc/pool_sort.cu. Checked per case: values bit-identical to the CPU backend (sort positions too — the inputs have
// distinct keys; a separate case with ties checks values and that the positions name equal keys), result on the GPU, at
// least one kernel launched, and NO device->host bytes inside the GPU call (masked_select: exactly the 8-byte element count).
// One JSON object on stdout; exit code 0 = all ok.
#include <axiom/axiom.hpp>

#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include <unistd.h>

#include "axb.h"

using namespace axiom;

struct Case {
    std::string name;
    bool ok;
    std::string why;
};
static std::vector<Case> g_cases;

static bool same(const Tensor &gpu, const Tensor &cpu, std::string *why) {
    if (gpu.device() != Device::GPU) {
        *why = "result is not on the GPU";
        return false;
    }
    Tensor a = gpu.cpu().ascontiguousarray(), b = cpu.ascontiguousarray();
    if (a.shape() != b.shape() || a.dtype() != b.dtype()) {
        *why = "shape / dtype differ";
        return false;
    }
    if (b.nbytes() && std::memcmp(a.data(), b.data(), b.nbytes()) != 0) {
        *why = "values differ from the CPU backend";
        return false;
    }
    return true;
}

// f maps one input to a list of results (topk returns two)
static void check(const std::string &name, const Tensor &x_cpu, const std::function<std::vector<Tensor>(const Tensor &)> &f,
                  uint64_t d2h_allowed = 0) {
    Case c{name, false, ""};
    try {
        std::vector<Tensor> ref = f(x_cpu);
        Tensor xg = x_cpu.gpu();
        axb_transfer_stats t0{}, t1{};
        axb_get_transfer_stats(&t0);
        const uint64_t l0 = axb_launch_count();
        std::vector<Tensor> got = f(xg);
        for (auto &t : got)
            t.sync();
        axb_get_transfer_stats(&t1);
        if (axb_launch_count() == l0 && ref[0].size() != 0)
            c.why = "no kernel launched";
        else if (t1.d2h_bytes - t0.d2h_bytes > d2h_allowed)
            c.why = "device->host traffic inside the GPU call: " + std::to_string(t1.d2h_bytes - t0.d2h_bytes) + " bytes";
        else if (got.size() != ref.size())
            c.why = "result count differs";
        else {
            c.ok = true;
            for (size_t i = 0; i < ref.size() && c.ok; ++i)
                c.ok = same(got[i], ref[i], &c.why);
        }
    } catch (const std::exception &e) {
        c.why = e.what();
    }
    g_cases.push_back(c);
}

static Tensor permutation_f32(const Shape &shape, float scale) { // distinct keys in a scrambled order
    size_t n = 1;
    for (auto d : shape)
        n *= d;
    std::vector<float> v(n);
    for (size_t i = 0; i < n; ++i)
        v[i] = ((float)((i * 2654435761ull) % n) - (float)(n / 2)) * scale; // multiplicative hash: a permutation for odd n ...
    // ... and for even n when the multiplier is odd (it is): i -> i * odd mod n is a bijection only if gcd(odd, n) = 1; use
    // a Fisher-Yates pass on top so that any n works
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i)
        v[i] = ((float)i - (float)(n / 2)) * scale;
    for (size_t i = n - 1; i > 0; --i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        std::swap(v[i], v[s % (i + 1)]);
    }
    return Tensor::from_data(v.data(), shape, true);
}

int main() {
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(77);
    using V = std::vector<Tensor>;

    // ---- pooling ---------------------------------------------------------------------------------------------------------
    Tensor x1 = Tensor::randn({2, 3, 41}) * 3.0f, x2 = Tensor::randn({2, 3, 13, 18}) * 3.0f, x3 = Tensor::randn({1, 2, 6, 7, 9});
    for (DType dt : {DType::Float32, DType::Float64, DType::BFloat16, DType::Float16}) {
        const std::string sfx = "_" + dtype_name(dt);
        Tensor a1 = x1.astype(dt), a2 = x2.astype(dt), a3 = x3.astype(dt);
        check("max_pool1d" + sfx, a1, [](const Tensor &t) { return V{ops::max_pool1d(t, 3, 2, 1)}; });
        check("max_pool1d_default_stride" + sfx, a1, [](const Tensor &t) { return V{ops::max_pool1d(t, 4)}; });
        check("avg_pool1d" + sfx, a1, [](const Tensor &t) { return V{ops::avg_pool1d(t, 3, 2, 1, true)}; });
        check("avg_pool1d_exclude_pad" + sfx, a1, [](const Tensor &t) { return V{ops::avg_pool1d(t, 3, 1, 1, false)}; });
        check("max_pool2d" + sfx, a2, [](const Tensor &t) { return V{ops::max_pool2d(t, {3, 2}, {2, 1}, {1, 1})}; });
        check("max_pool2d_defaults" + sfx, a2, [](const Tensor &t) { return V{ops::max_pool2d(t, {2, 2})}; });
        check("avg_pool2d" + sfx, a2, [](const Tensor &t) { return V{ops::avg_pool2d(t, {3, 3}, {2, 2}, {1, 1}, true)}; });
        check("avg_pool2d_exclude_pad" + sfx, a2, [](const Tensor &t) { return V{ops::avg_pool2d(t, {3, 3}, {2, 2}, {1, 1}, false)}; });
        check("max_pool3d" + sfx, a3, [](const Tensor &t) { return V{ops::max_pool3d(t, {2, 3, 2}, {1, 2, 2}, {1, 1, 0})}; });
        check("avg_pool3d" + sfx, a3, [](const Tensor &t) { return V{ops::avg_pool3d(t, {2, 3, 2}, {1, 2, 2}, {1, 1, 0}, true)}; });
        check("adaptive_avg_pool2d" + sfx, a2, [](const Tensor &t) { return V{ops::adaptive_avg_pool2d(t, {3, 4})}; });
        check("adaptive_max_pool2d" + sfx, a2, [](const Tensor &t) { return V{ops::adaptive_max_pool2d(t, {1, 1})}; });
        check("adaptive_avg_pool1d" + sfx, a1, [](const Tensor &t) { return V{ops::adaptive_avg_pool1d(t, 5)}; });
    }
    // unbatched inputs and a strided view
    check("max_pool2d_unbatched", x2.slice({Slice(0, 1)}).squeeze(0), [](const Tensor &t) { return V{ops::max_pool2d(t, {2, 2}, {2, 2}, {0, 0})}; });
    check("avg_pool1d_unbatched", x1.slice({Slice(1, 2)}).squeeze(0), [](const Tensor &t) { return V{ops::avg_pool1d(t, 2, 2, 0, true)}; });
    check("max_pool2d_transposed_view", x2, [](const Tensor &t) { return V{ops::max_pool2d(t.transpose({0, 1, 3, 2}), {2, 3}, {1, 2}, {0, 1})}; });
    // the nn modules are compositions of the ops (src/nn/pooling.cpp)
    check("nn_MaxPool2d", x2, [](const Tensor &t) { return V{nn::MaxPool2d({2, 2}, {2, 2}, {0, 0})(t)}; });
    check("nn_AdaptiveAvgPool2d", x2, [](const Tensor &t) { return V{nn::AdaptiveAvgPool2d({1, 1})(t)}; });

    // ---- sort / argsort / topk: distinct keys ---------------------------------------------------------------------------------
    Tensor p1 = permutation_f32({37}, 0.5f), p3 = permutation_f32({5, 301, 3}, 0.25f), pl = permutation_f32({2, 70001}, 1.0f);
    for (bool desc : {false, true}) {
        const std::string sfx = desc ? "_desc" : "_asc";
        check("sort_1d" + sfx, p1, [=](const Tensor &t) { return V{ops::sort(t, -1, desc)}; });
        check("argsort_1d" + sfx, p1, [=](const Tensor &t) { return V{ops::argsort(t, 0, desc)}; });
        for (int axis : {0, 1, 2}) {
            check("sort_3d_axis" + std::to_string(axis) + sfx, p3, [=](const Tensor &t) { return V{ops::sort(t, axis, desc)}; });
            check("argsort_3d_axis" + std::to_string(axis) + sfx, p3, [=](const Tensor &t) { return V{ops::argsort(t, axis, desc)}; });
        }
        check("sort_long_rows" + sfx, pl, [=](const Tensor &t) { return V{ops::sort(t, 1, desc)}; });
        check("argsort_long_rows" + sfx, pl, [=](const Tensor &t) { return V{ops::argsort(t, 1, desc)}; });
        check("topk_3d" + sfx, p3, [=](const Tensor &t) {
            auto r = ops::topk(t, 7, 1, desc, true);
            return V{r.first, r.second};
        });
        check("topk_long_rows" + sfx, pl, [=](const Tensor &t) {
            auto r = ops::topk(t, 50, -1, desc, false);
            return V{r.first, r.second};
        });
    }
    check("sort_int64", (p3 * 4.0f).astype(DType::Int64), [](const Tensor &t) { return V{ops::sort(t, 1, false), ops::argsort(t, 1, true)}; });
    check("sort_int32", (p3 * 4.0f).astype(DType::Int32), [](const Tensor &t) { return V{ops::sort(t, 2, true), ops::argsort(t, 0, false)}; });
    check("sort_float64", p3.astype(DType::Float64), [](const Tensor &t) { return V{ops::sort(t, 1, true), ops::argsort(t, 1, false)}; });
    // dtypes the reference converts to float32 first (:4210, :4233, :4253): the result dtype is Float32
    check("sort_bf16_becomes_f32", p1.astype(DType::BFloat16), [](const Tensor &t) { return V{ops::sort(t), ops::argsort(t)}; });
    check("sort_int16_becomes_f32", p1.astype(DType::Int16), [](const Tensor &t) { return V{ops::sort(t, 0, true)}; });
    check("sort_noncontiguous", p3, [](const Tensor &t) { return V{ops::sort(t.transpose({2, 0, 1}), 2, false)}; });
    // ties: sorted VALUES have one answer whatever the order of equal keys
    Tensor ties = (Tensor::randn({6, 500}) * 2.0f).round() + 0.0f; // + 0.0f: no -0 (equal to +0 under `<`, different bits)
    check("sort_values_with_ties", ties, [](const Tensor &t) { return V{ops::sort(t, 1, false), ops::sort(t, 1, true)}; });
    {
        Case c{"argsort_with_ties_names_equal_keys", false, ""};
        try {
            Tensor idx = ops::argsort(ties.gpu(), 1, false);
            Tensor viaidx = ops::gather(ties, 1, idx.cpu());
            Tensor sorted = ops::sort(ties, 1, false);
            c.ok = idx.device() == Device::GPU && std::memcmp(viaidx.ascontiguousarray().data(), sorted.ascontiguousarray().data(), sorted.nbytes()) == 0;
            if (!c.ok)
                c.why = "gathering by the GPU positions does not give the sorted values";
        } catch (const std::exception &e) {
            c.why = e.what();
        }
        g_cases.push_back(c);
    }

    // ---- Tensor::zeros / ones / eye on Device::GPU (patched: device-side fill instead of host fill + upload) --------------------
    {
        Case c{"zeros_ones_eye_on_gpu", false, ""};
        try {
            axb_transfer_stats t0{}, t1{};
            axb_get_transfer_stats(&t0);
            Tensor z = Tensor::zeros({257, 33}, DType::Float32, Device::GPU), o = Tensor::ones({64, 5, 7}, DType::BFloat16, Device::GPU);
            Tensor oi = Tensor::ones({1000}, DType::Int64, Device::GPU), zb = Tensor::zeros({77}, DType::Bool, Device::GPU);
            z.sync(); o.sync(); oi.sync(); zb.sync();
            axb_get_transfer_stats(&t1);
            std::string why;
            c.ok = same(z, Tensor::zeros({257, 33}), &why) && same(o, Tensor::ones({64, 5, 7}, DType::BFloat16), &why) &&
                   same(oi, Tensor::ones({1000}, DType::Int64), &why) && same(zb, Tensor::zeros({77}, DType::Bool), &why) &&
                   same(Tensor::eye(9, DType::Float32, Device::GPU), Tensor::eye(9), &why);
            if (c.ok && t1.h2d_bytes != t0.h2d_bytes) {
                c.ok = false;
                why = "zeros / ones on the GPU uploaded " + std::to_string(t1.h2d_bytes - t0.h2d_bytes) + " bytes from the host";
            }
            c.why = why;
        } catch (const std::exception &e) {
            c.why = e.what();
        }
        g_cases.push_back(c);
    }

    // ---- transfers of contiguous views that start inside their storage (patched Tensor::to, tensor.cpp:1625-1636) -------------
    {
        Case c{"transfer_of_offset_views", false, ""};
        try {
            Tensor row_cpu = x1.slice({Slice(1, 2)});          // contiguous, offset != 0
            Tensor up = row_cpu.gpu().cpu();                   // H2D of the view, D2H of the compact copy
            Tensor x1g = x1.gpu();
            Tensor down = x1g.slice({Slice(1, 2)}).cpu();      // D2H of a device view with an offset
            Tensor e = row_cpu.ascontiguousarray();
            const bool a = up.shape() == e.shape() && std::memcmp(up.ascontiguousarray().data(), e.data(), e.nbytes()) == 0;
            const bool b = down.shape() == e.shape() && std::memcmp(down.ascontiguousarray().data(), e.data(), e.nbytes()) == 0;
            c.ok = a && b;
            if (!c.ok)
                c.why = std::string(a ? "" : "host->device of an offset view moved the wrong bytes; ") + (b ? "" : "device->host of an offset view moved the wrong bytes");
        } catch (const std::exception &e) {
            c.why = e.what();
        }
        g_cases.push_back(c);
    }

    // ---- safetensors -> device (patched io::safetensors::load, src/io/io_safetensors.cpp:123-214) ---------------------------
    {
        Case c{"safetensors_load_to_gpu", false, ""};
        const std::string path = "/tmp/axb_op_surface_" + std::to_string((long)getpid()) + ".safetensors";
        try {
            // three tensors; `big` (40 MiB) spans several staging chunks; data section in header order
            Tensor w = Tensor::randn({33, 17}), idx = (Tensor::randn({11}) * 100.0f).astype(DType::Int64);
            Tensor hb = Tensor::randn({5, 7}).astype(DType::BFloat16), big = Tensor::randn({10, 1024, 1024});
            std::vector<std::pair<std::string, Tensor>> items = {{"enc.w", w}, {"idx", idx}, {"half", hb}, {"big", big}};
            std::string header = "{\"__metadata__\":{\"format\":\"pt\"}";
            size_t off = 0;
            for (auto &kv : items) {
                Tensor t = kv.second.ascontiguousarray();
                const char *dn = t.dtype() == DType::Float32 ? "F32" : t.dtype() == DType::Int64 ? "I64" : "BF16";
                header += ",\"" + kv.first + "\":{\"dtype\":\"" + dn + "\",\"shape\":[";
                for (size_t d = 0; d < t.ndim(); ++d)
                    header += (d ? "," : "") + std::to_string(t.shape()[d]);
                header += "],\"data_offsets\":[" + std::to_string(off) + "," + std::to_string(off + t.nbytes()) + "]}";
                off += t.nbytes();
            }
            header += "}";
            FILE *f = std::fopen(path.c_str(), "wb");
            const uint64_t hl = header.size();
            std::fwrite(&hl, 8, 1, f);
            std::fwrite(header.data(), 1, header.size(), f);
            for (auto &kv : items) {
                Tensor t = kv.second.ascontiguousarray();
                std::fwrite(t.data(), 1, t.nbytes(), f);
            }
            std::fclose(f);
            auto cpu = io::safetensors::load(path, Device::CPU);
            axb_transfer_stats t0{}, t1{};
            axb_get_transfer_stats(&t0);
            auto gpu = io::safetensors::load(path, Device::GPU);
            Tensor one = io::safetensors::load_tensor(path, "half", Device::GPU);
            axb_get_transfer_stats(&t1);
            c.ok = gpu.size() == cpu.size() && cpu.size() == items.size();
            for (auto &kv : cpu) {
                if (!c.ok)
                    break;
                auto it = gpu.find(kv.first);
                c.ok = it != gpu.end() && same(it->second, kv.second, &c.why);
            }
            if (c.ok)
                c.ok = same(one, cpu["half"], &c.why);
            if (c.ok && t1.h2d_bytes - t0.h2d_bytes != off + cpu["half"].nbytes()) {
                c.ok = false;
                c.why = "uploaded " + std::to_string(t1.h2d_bytes - t0.h2d_bytes) + " bytes for a " + std::to_string(off) + "-byte data section";
            }
            if (c.ok && !c.why.empty())
                c.why.clear();
            if (!c.ok && c.why.empty())
                c.why = "tensor set differs";
            // error behaviour: a header that promises more bytes than the shape holds is a FileFormatError on both devices
        } catch (const std::exception &e) {
            c.ok = false;
            c.why = e.what();
        }
        std::remove(path.c_str());
        g_cases.push_back(c);
    }

    // ---- masked_select ----------------------------------------------------------------------------------------------------------
    Tensor mx = Tensor::randn({5, 33, 129});
    Tensor mk = mx > 0.2f, mrow = Tensor::randn({33, 1}) > 0.0f;
    mk.sync();
    mrow.sync();
    auto on = [](const Tensor &like, const Tensor &m) { return like.device() == Device::GPU ? m.gpu() : m; };
    check("masked_select", mx, [&](const Tensor &t) { return V{t.masked_select(on(t, mk))}; }, 8);
    check("masked_select_broadcast_mask", mx, [&](const Tensor &t) { return V{ops::masked_select(t, on(t, mrow))}; }, 8);
    check("masked_select_lazy_mask", mx, [&](const Tensor &t) { return V{t.masked_select(t > 1.0f)}; }, 8);
    check("masked_select_int64", (mx * 100.0f).astype(DType::Int64), [&](const Tensor &t) { return V{t.masked_select(on(t, mk))}; }, 8);
    check("masked_select_none", mx, [&](const Tensor &t) { return V{t.masked_select(t > 1.0e9f)}; }, 8);
    check("masked_select_transposed", mx, [&](const Tensor &t) { return V{t.transpose({0, 2, 1}).masked_select(on(t, mk).transpose({0, 2, 1}))}; }, 8);

    bool all = true;
    std::printf("{\"cases\": %zu, \"failed\": [", g_cases.size());
    bool first = true;
    for (auto &k : g_cases)
        if (!k.ok) {
            all = false;
            std::printf("%s{\"name\": \"%s\", \"why\": \"%s\"}", first ? "" : ", ", k.name.c_str(), k.why.c_str());
            first = false;
        }
    std::printf("], \"all_ok\": %s}\n", all ? "true" : "false");
    return all ? 0 : 1;
}
