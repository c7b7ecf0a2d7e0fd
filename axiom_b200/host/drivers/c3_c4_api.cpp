// c3_c4_api — BASELINE configs 3 and 4 driven through the reference's own Tensor API on Device::GPU at the BASELINE sizes
// ("x.gpu().softmax(-1), ops::layer_norm, x.argmax(1) and bf16 matmul reach the wide, cluster and pair
// kernels through cuda_graph.cpp" was untested). For every call it checks
//   * NO device->host traffic between the call and its materialisation (axb_get_transfer_stats): a silent drop to the
//     reference's CPU path (src/tensor/operations.cpp:913-917, src/graph/graph_executor.cpp:346-349) would show up here,
//   * kernels were launched (axb_launch_count) and the EXPECTED kernel family ran (axb_rownorm_last_path /
//     axb_matmul_last_path),
//   * parity against the same call on Device::CPU of the same library: fp32 at full size (sum 1e-5 * sum|x|, max and
//     argmax bit-exact, softmax / layer_norm 1e-4 relative to the reference + 1e-5 against an fp64 row check), bf16 on
//     the first `half_rows` rows (the CPU bf16 paths are scalar), C4 on rows of two of the 64 matrices against fp32 sgemm
//     of the same bf16 inputs (<= 1e-2 relative, north_star).
// Device time per call is measured with events on the backend stream. One JSON object on stdout; exit code 0 = all ok.
// usage: c3_c4_api [rows cols [batch m [iters]]]       defaults 8192 32768 64 4096 5
#include <axiom/axiom.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "axb.h"
#include "backends/metal/metal_common.hpp"

using namespace axiom;

static void gpu_sync() { backends::metal::MetalExecutionStream::instance().synchronize(); }

// deterministic N(0,1)-like fill (Box-Muller over a counter hash), parallel: Tensor::randn is a serial host loop
// (src/tensor/tensor.cpp:1884-1904: 12.9 s for 268 M elements)
static inline uint32_t mix(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return (uint32_t)x;
}
static void fill_normal(Tensor &t, uint64_t seed, float scale = 1.0f) {
    float *p = t.typed_data<float>();
    const int64_t n = (int64_t)t.size();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i += 2) {
        const float u1 = (mix(seed * 0x9E3779B97F4A7C15ull + i) + 1.0f) * (1.0f / 4294967808.0f);
        const float u2 = mix(seed * 0xD1B54A32D192ED03ull + i + 1) * (1.0f / 4294967296.0f);
        const float r = std::sqrt(-2.0f * std::log(u1)) * scale;
        p[i] = r * std::cos(6.2831853f * u2);
        if (i + 1 < n)
            p[i + 1] = r * std::sin(6.2831853f * u2);
    }
}

struct Probe {
    axb_transfer_stats x0{};
    uint64_t l0 = 0;
    void begin() {
        gpu_sync();
        axb_get_transfer_stats(&x0);
        l0 = axb_launch_count();
    }
    // returns launches; fills d2h / h2d deltas
    uint64_t end(uint64_t *d2h, uint64_t *h2d) {
        gpu_sync();
        axb_transfer_stats x1{};
        axb_get_transfer_stats(&x1);
        *d2h = x1.d2h_bytes - x0.d2h_bytes;
        *h2d = x1.h2d_bytes - x0.h2d_bytes;
        return axb_launch_count() - l0;
    }
};

struct Result {
    std::string name;
    bool ok = false;
    uint64_t launches = 0, d2h = 0, h2d = 0;
    int path = -1;
    double ms = 0, metric = 0; // GB/s or TFLOP/s
    double err = 0;
    std::string note;
};

static double time_gpu(const std::function<Tensor()> &f, int iters) {
    for (int i = 0; i < 2; ++i)
        f().sync();
    axb_event_t e0, e1;
    axb_event_create(&e0);
    axb_event_create(&e1);
    gpu_sync();
    axb_event_record(e0, nullptr);
    for (int i = 0; i < iters; ++i)
        f().sync();
    axb_event_record(e1, nullptr);
    axb_event_sync(e1);
    float ms = 0;
    axb_event_elapsed_ms(e0, e1, &ms);
    axb_event_destroy(e0);
    axb_event_destroy(e1);
    return ms / iters;
}

int main(int argc, char **argv) {
    const size_t R = argc > 2 ? std::strtoul(argv[1], nullptr, 10) : 8192;
    const size_t N = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 32768;
    const size_t NB = argc > 4 ? std::strtoul(argv[3], nullptr, 10) : 64;
    const size_t M = argc > 4 ? std::strtoul(argv[4], nullptr, 10) : 4096;
    const int iters = argc > 5 ? std::atoi(argv[5]) : 5;
    const size_t half_rows = std::min<size_t>(R, 256);
    ops::OperationRegistry::initialize_builtin_operations();
    std::vector<Result> results;
    bool all_ok = true;

    // ---- config 3, fp32 --------------------------------------------------------------------------------------------
    Tensor x({R, N}, DType::Float32, Device::CPU);
    fill_normal(x, 7);
    Tensor w({N}, DType::Float32, Device::CPU), b({N}, DType::Float32, Device::CPU);
    fill_normal(w, 8, 0.1f);
    fill_normal(b, 9, 0.1f);
    {
        float *wp = w.typed_data<float>();
        for (size_t i = 0; i < N; ++i)
            wp[i] += 1.0f;
    }
    Tensor xg = x.gpu(), wg = w.gpu(), bg = b.gpu();
    const double in_b = (double)R * N * 4;

    auto run = [&](const std::string &name, const std::function<Tensor(const Tensor &)> &op, const Tensor &gin, const Tensor &cin,
                   double bytes, int want_path, const std::function<double(const Tensor &, const Tensor &)> &compare, double tol) {
        Result r;
        r.name = name;
        Probe pr;
        pr.begin();
        Tensor y = op(gin);
        y.sync();
        r.launches = pr.end(&r.d2h, &r.h2d);
        r.path = want_path >= 0 ? axb_rownorm_last_path() : -1;
        r.ms = time_gpu([&] { return op(gin); }, iters);
        r.metric = bytes / r.ms / 1e6;
        Tensor ref = op(cin);
        ref.sync();
        Tensor got = y.cpu();
        r.err = (got.shape() == ref.shape() && got.dtype() == ref.dtype()) ? compare(got, ref) : 1e30;
        r.ok = y.device() == Device::GPU && r.d2h == 0 && r.launches >= 1 && r.err <= tol && (want_path < 0 || r.path == want_path);
        if (r.d2h != 0)
            r.note += "device->host traffic inside a GPU call (CPU fallback?) ";
        if (want_path >= 0 && r.path != want_path)
            r.note += "unexpected kernel family ";
        all_ok = all_ok && r.ok;
        results.push_back(r);
    };

    // |got - ref| <= tol * sum|x| per row  -> returns max over rows of |diff| / sum|x|
    std::vector<double> row_abs(R);
    {
        const float *xp = x.typed_data<float>();
#pragma omp parallel for
        for (int64_t i = 0; i < (int64_t)R; ++i) {
            double s = 0;
            for (size_t j = 0; j < N; ++j)
                s += std::fabs((double)xp[i * N + j]);
            row_abs[i] = s;
        }
    }
    auto cmp_sum = [&](const Tensor &g, const Tensor &r) {
        double worst = 0;
        for (size_t i = 0; i < g.size(); ++i)
            worst = std::max(worst, std::fabs((double)g.typed_data<float>()[i] - r.typed_data<float>()[i]) / row_abs[i]);
        return worst;
    };
    auto cmp_bits = [&](const Tensor &g, const Tensor &r) {
        return std::memcmp(g.data(), r.data(), r.nbytes()) == 0 ? 0.0 : 1.0;
    };
    auto cmp_rel = [&](const Tensor &g, const Tensor &r) { // max |d| / (|ref| + floor)
        const float *gp = g.typed_data<float>(), *rp = r.typed_data<float>();
        double worst = 0;
        const int64_t n = (int64_t)r.size();
#pragma omp parallel for reduction(max : worst)
        for (int64_t i = 0; i < n; ++i) {
            const double d = std::fabs((double)gp[i] - rp[i]) / (std::fabs((double)rp[i]) + 1e-7);
            if (d > worst)
                worst = d;
        }
        return worst;
    };
    auto cmp_abs_rel = [&](double atol) {
        return [atol](const Tensor &g, const Tensor &r) {
            const float *gp = g.typed_data<float>(), *rp = r.typed_data<float>();
            double worst = 0;
            const int64_t n = (int64_t)r.size();
#pragma omp parallel for reduction(max : worst)
            for (int64_t i = 0; i < n; ++i) {
                const double d = std::fabs((double)gp[i] - rp[i]) / (std::fabs((double)rp[i]) + atol);
                if (d > worst)
                    worst = d;
            }
            return worst;
        };
    };

    run("c3_sum_f32", [](const Tensor &t) { return t.sum({1}); }, xg, x, in_b + R * 4, -1, cmp_sum, 1e-5);
    run("c3_max_f32", [](const Tensor &t) { return t.max(1); }, xg, x, in_b + R * 4, -1, cmp_bits, 0);
    run("c3_argmax_f32", [](const Tensor &t) { return t.argmax(1); }, xg, x, in_b + R * 8, -1, cmp_bits, 0);
    // softmax: the reference's serial fp32 denominator is the less accurate side beyond n = 4096 (tests/test_reduce_gpu.py)
    run("c3_softmax_f32", [](const Tensor &t) { return t.softmax(-1); }, xg, x, 2 * in_b, N * 4 >= 32768 ? 1 : -1, cmp_rel, 1e-4);
    run("c3_layer_norm_f32", [&](const Tensor &t) { return ops::layer_norm(t, t.device() == Device::GPU ? wg : w, t.device() == Device::GPU ? bg : b, -1, 1e-5f); },
        xg, x, 2 * in_b + 2 * N * 4, N >= 8192 ? 2 : -1, cmp_abs_rel(1e-2), 1e-4);

    // ---- config 3, bf16 (GPU at full size; CPU reference on the first half_rows rows) -------------------------------
    {
        Tensor xbg = xg.astype(DType::BFloat16);
        xbg.sync();
        Tensor xb_head = xbg.slice({Slice(0, (int64_t)half_rows)}).cpu(); // the SAME bf16 bits the GPU ops read
        const double hb = (double)R * N * 2;
        auto head = [&](const Tensor &t) { return t.slice({Slice(0, (int64_t)half_rows)}); };
        auto run_bf = [&](const std::string &name, const std::function<Tensor(const Tensor &)> &op, double bytes, int want_path, double tol,
                          bool exact) {
            Result r;
            r.name = name;
            Probe pr;
            pr.begin();
            Tensor y = op(xbg);
            y.sync();
            r.launches = pr.end(&r.d2h, &r.h2d);
            r.path = want_path >= 0 ? axb_rownorm_last_path() : -1;
            r.ms = time_gpu([&] { return op(xbg); }, iters);
            r.metric = bytes / r.ms / 1e6;
            Tensor ref = op(xb_head);
            ref.sync();
            Tensor got = head(y).cpu();
            if (exact) {
                r.err = (got.nbytes() == ref.nbytes() && std::memcmp(got.ascontiguousarray().data(), ref.data(), ref.nbytes()) == 0) ? 0 : 1;
            } else {
                Tensor gf = got.astype(DType::Float32).ascontiguousarray(), rf = ref.astype(DType::Float32).ascontiguousarray();
                r.err = cmp_abs_rel(1e-2)(gf, rf);
            }
            r.ok = y.device() == Device::GPU && r.d2h == 0 && r.launches >= 1 && r.err <= tol && (want_path < 0 || r.path == want_path);
            if (r.d2h != 0)
                r.note += "device->host traffic inside a GPU call (CPU fallback?) ";
            if (want_path >= 0 && r.path != want_path)
                r.note += "unexpected kernel family ";
            all_ok = all_ok && r.ok;
            results.push_back(r);
        };
        // bf16 tolerances: one rounding to bf16 at the end of either side (2^-8 relative), sums differ in order only
        run_bf("c3_sum_bf16", [](const Tensor &t) { return t.sum({1}); }, hb + R * 2, -1, 1.6e-2, false);
        run_bf("c3_max_bf16", [](const Tensor &t) { return t.max(1); }, hb + R * 2, -1, 0, true);
        run_bf("c3_argmax_bf16", [](const Tensor &t) { return t.argmax(1); }, hb + R * 8, -1, 0, true);
        run_bf("c3_softmax_bf16", [](const Tensor &t) { return t.softmax(-1); }, 2 * hb, N * 2 >= 32768 ? 1 : -1, 1.6e-2, false);
        run_bf("c3_layer_norm_bf16",
               [&](const Tensor &t) { return ops::layer_norm(t, t.device() == Device::GPU ? wg : w, t.device() == Device::GPU ? bg : b, -1, 1e-5f); },
               2 * hb + 2 * N * 4, N >= 8192 ? 2 : -1, 3.2e-2, false);
    }
    xg = Tensor();
    x = Tensor();

    // ---- config 4: bf16 batched matmul {NB, M, M} x {NB, M, M} -------------------------------------------------------
    {
        Result r;
        r.name = "c4_bmm_bf16";
        Tensor a0({1, M, M}, DType::Float32, Device::CPU), b0({1, M, M}, DType::Float32, Device::CPU);
        fill_normal(a0, 11);
        fill_normal(b0, 12, 1.0f / 64.0f);
        Tensor sa({NB, 1, 1}, DType::Float32, Device::CPU), sb({NB, 1, 1}, DType::Float32, Device::CPU);
        for (size_t i = 0; i < NB; ++i) { // distinct matrices per batch entry, built ON the device by one fused kernel each
            sa.typed_data<float>()[i] = 1.0f + 0.01f * (float)i;
            sb.typed_data<float>()[i] = 1.0f - 0.005f * (float)i;
        }
        Tensor A = (a0.gpu() * sa.gpu()).astype(DType::BFloat16), B = (b0.gpu() * sb.gpu()).astype(DType::BFloat16);
        A.sync();
        B.sync();
        Probe pr;
        pr.begin();
        Tensor C = A.matmul(B);
        C.sync();
        r.launches = pr.end(&r.d2h, &r.h2d);
        r.path = axb_matmul_last_path();
        r.ms = time_gpu([&] { return A.matmul(B); }, std::max(2, iters / 2));
        r.metric = 2.0 * NB * M * M * M / r.ms / 1e9; // TFLOP/s
        // parity: 128 rows of the first and the last matrix against fp32 sgemm of the same bf16 inputs
        double worst = 0;
        const size_t rows = std::min<size_t>(128, M);
        for (size_t bi : {(size_t)0, NB - 1}) {
            // NB: Tensor::to() copies a CONTIGUOUS view from the START of its storage (tensor.cpp:1626-1634 ignores offset_),
            // so a slice of matrix 63 must become a fresh tensor on the device (Cast op) before it travels to the host
            Tensor As = A.slice({Slice((int64_t)bi, (int64_t)bi + 1), Slice(0, (int64_t)rows)}).astype(DType::Float32).cpu();
            Tensor Bs = B.slice({Slice((int64_t)bi, (int64_t)bi + 1)}).astype(DType::Float32).cpu();
            Tensor ref = As.matmul(Bs).ascontiguousarray();
            Tensor got = C.slice({Slice((int64_t)bi, (int64_t)bi + 1), Slice(0, (int64_t)rows)}).astype(DType::Float32).cpu().ascontiguousarray();
            // norm-wise bound per element: |d| <= 1e-2 * (|ref| + row scale); the row scale guards cancellation
            const float *gp = got.typed_data<float>(), *rp = ref.typed_data<float>();
            double scale = 0;
            for (size_t i = 0; i < ref.size(); ++i)
                scale += std::fabs((double)rp[i]);
            scale /= (double)ref.size();
            for (size_t i = 0; i < ref.size(); ++i)
                worst = std::max(worst, std::fabs((double)gp[i] - rp[i]) / (std::fabs((double)rp[i]) + scale));
        }
        r.err = worst;
        const bool big = NB * ((M + 255) / 256) * ((M + 255) / 256) >= 148;
        r.ok = C.device() == Device::GPU && C.dtype() == DType::BFloat16 && r.d2h == 0 && r.launches >= 1 && r.err <= 1e-2 &&
               (!big || r.path == 2);
        if (big && r.path != 2)
            r.note += "expected the CTA-pair tcgen05 kernel ";
        all_ok = all_ok && r.ok;
        results.push_back(r);
    }

    std::printf("{\"rows\": %zu, \"cols\": %zu, \"batch\": %zu, \"m\": %zu, \"all_ok\": %s, \"calls\": {", R, N, NB, M, all_ok ? "true" : "false");
    for (size_t i = 0; i < results.size(); ++i) {
        const Result &r = results[i];
        std::printf("%s\"%s\": {\"ok\": %s, \"gpu_launches\": %llu, \"d2h_bytes\": %llu, \"h2d_bytes\": %llu, \"kernel_path\": %d, \"gpu_ms\": %.4f, "
                    "\"%s\": %.1f, \"max_err\": %.3g, \"note\": \"%s\"}",
                    i ? ", " : "", r.name.c_str(), r.ok ? "true" : "false", (unsigned long long)r.launches, (unsigned long long)r.d2h,
                    (unsigned long long)r.h2d, r.path, r.ms, r.name[1] == '4' ? "TFLOPs" : "GBps", r.metric, r.err, r.note.c_str());
    }
    std::printf("}}\n");
    return all_ok ? 0 : 1;
}
