// axiom_chains — BASELINE configs 1 and 2 written exactly as a user of the reference writes them (Tensor API, lazy
// graph), run on Device::CPU and Device::GPU of the SAME unmodified library. Prints one JSON object: parity of the GPU
// result against the CPU backend, kernels launched per materialisation, and wall-clock per call on both devices.
//   config 1: x.matmul(x.T()).softmax(-1).sum({1})                      x fp32 {1024,1024}
//   config 2: ((a.relu() + 1.0f) * b).where(c > 0, 0)                   a {R,1}, b {1,C}, c {R,C} — the LITERAL BASELINE
//             expression: scalar operands 1.0f, int 0 (comparison) and int 0 (where), nothing pre-uploaded
//   config 2': the same chain with the where-constant passed as a pre-uploaded {1} tensor (round-1 form, kept for comparison)
// usage: axiom_chains [rows cols [iters]]
#include <axiom/axiom.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "axb.h"
#include "backends/metal/metal_common.hpp"

using namespace axiom;
using Clock = std::chrono::steady_clock;

static void gpu_sync() { backends::metal::MetalExecutionStream::instance().synchronize(); }
static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

int main(int argc, char **argv) {
    const size_t R = argc > 2 ? std::strtoul(argv[1], nullptr, 10) : 4096;
    const size_t C = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 16384;
    const int iters = argc > 3 ? std::atoi(argv[3]) : 10;
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(1234);

    // ---- config 1 ------------------------------------------------------------------------------------------------
    auto x = Tensor::randn({1024, 1024}, DType::Float32, Device::CPU);
    auto c1 = [](const Tensor &t) { return t.matmul(t.T()).softmax(-1).sum({1}); };
    Tensor r1_cpu = c1(x);
    r1_cpu.sync();
    auto t0 = Clock::now();
    for (int i = 0; i < 3; ++i)
        c1(x).sync();
    const double c1_cpu_ms = ms_since(t0) / 3;
    Tensor xg = x.gpu();
    Tensor r1_gpu = c1(xg);
    r1_gpu.sync();
    gpu_sync();
    const uint64_t l0 = axb_launch_count();
    c1(xg).sync();
    const uint64_t c1_launches = axb_launch_count() - l0;
    gpu_sync();
    t0 = Clock::now();
    for (int i = 0; i < iters; ++i)
        c1(xg).sync();
    gpu_sync();
    const double c1_gpu_ms = ms_since(t0) / iters;
    const bool c1_ok = r1_gpu.device() == Device::GPU && r1_cpu.allclose(r1_gpu.cpu(), 1e-5, 1e-6);

    // ---- config 2 ------------------------------------------------------------------------------------------------
    auto a = Tensor::randn({R, 1}), b = Tensor::randn({1, C}), c = Tensor::randn({R, C});
    auto c2 = [](const Tensor &a, const Tensor &b, const Tensor &c) { return ((a.relu() + 1.0f) * b).where(c > 0, 0); };
    auto zero = Tensor::zeros({1});
    auto c2t = [](const Tensor &a, const Tensor &b, const Tensor &c, const Tensor &z) {
        return ((a.relu() + 1.0f) * b).where(c > 0.0f, z);
    };
    Tensor r2_cpu = c2(a, b, c);
    r2_cpu.sync();
    t0 = Clock::now();
    for (int i = 0; i < 3; ++i)
        c2(a, b, c).sync();
    const double c2_cpu_ms = ms_since(t0) / 3;
    Tensor ag = a.gpu(), bg = b.gpu(), cg = c.gpu(), zg = zero.gpu();
    Tensor r2_gpu = c2(ag, bg, cg);
    r2_gpu.sync();
    gpu_sync();
    axb_transfer_stats x0{}, x1{};
    axb_get_transfer_stats(&x0);
    const uint64_t l1 = axb_launch_count();
    c2(ag, bg, cg).sync();
    const uint64_t c2_launches = axb_launch_count() - l1;
    gpu_sync();
    axb_get_transfer_stats(&x1);
    // wall clock per call (graph construction + lowering + enqueue + the final sync amortised over `iters` calls) ...
    t0 = Clock::now();
    for (int i = 0; i < iters; ++i)
        c2(ag, bg, cg).sync();
    gpu_sync();
    const double c2_gpu_ms = ms_since(t0) / iters;
    // ... and device time between two events on the backend stream around the same calls
    axb_event_t e0, e1;
    axb_event_create(&e0);
    axb_event_create(&e1);
    axb_event_record(e0, nullptr);
    for (int i = 0; i < iters; ++i)
        c2(ag, bg, cg).sync();
    axb_event_record(e1, nullptr);
    axb_event_sync(e1);
    float c2_dev_ms = 0;
    axb_event_elapsed_ms(e0, e1, &c2_dev_ms);
    c2_dev_ms /= iters;
    // pre-uploaded-constant form
    c2t(ag, bg, cg, zg).sync();
    gpu_sync();
    t0 = Clock::now();
    for (int i = 0; i < iters; ++i)
        c2t(ag, bg, cg, zg).sync();
    gpu_sync();
    const double c2t_gpu_ms = ms_since(t0) / iters;
    // elementwise fp32 chain: bit-exact against the CPU backend
    Tensor r2_back = r2_gpu.cpu();
    const bool c2_ok = r2_gpu.device() == Device::GPU && r2_back.shape() == r2_cpu.shape() &&
                       std::memcmp(r2_back.data(), r2_cpu.data(), r2_cpu.nbytes()) == 0;
    const double c2_bytes = (double)R * C * 8 + R * 4 + C * 4 + 4;

    std::printf("{\"config1\": {\"parity_ok\": %s, \"gpu_launches\": %llu, \"gpu_ms\": %.4f, \"cpu_ms\": %.3f},"
                " \"config2\": {\"expression\": \"((a.relu()+1.0f)*b).where(c>0,0)\", \"rows\": %zu, \"cols\": %zu, \"bit_exact\": %s,"
                " \"gpu_launches\": %llu, \"h2d_bytes_per_call\": %llu, \"h2d_copies_per_call\": %llu, \"d2h_bytes_per_call\": %llu,"
                " \"gpu_ms\": %.4f, \"gpu_device_ms\": %.4f, \"gpu_GBps\": %.1f, \"preuploaded_constant_gpu_ms\": %.4f,"
                " \"cpu_ms\": %.3f, \"cpu_GBps\": %.2f}}\n",
                c1_ok ? "true" : "false", (unsigned long long)c1_launches, c1_gpu_ms, c1_cpu_ms, R, C, c2_ok ? "true" : "false",
                (unsigned long long)c2_launches, (unsigned long long)(x1.h2d_bytes - x0.h2d_bytes),
                (unsigned long long)(x1.h2d_calls - x0.h2d_calls), (unsigned long long)(x1.d2h_bytes - x0.d2h_bytes), c2_gpu_ms,
                (double)c2_dev_ms, c2_bytes / c2_gpu_ms / 1e6, c2t_gpu_ms, c2_cpu_ms, c2_bytes / c2_cpu_ms / 1e6);
    return (c1_ok && c2_ok) ? 0 : 1;
}
