// whisper_encoder — BASELINE config 5: the whisper_poc encoder forward (conv1d + attention + MLP) with synthetic
// random-init weights, through the reference's own model code. The model classes come from the UNMODIFIED example file
// (examples/whisper_poc.cpp is included as-is with its main() renamed); weights follow its random branch
// (whisper_poc.cpp:407-443). The same encoder runs on Device::CPU (one sample, the reference backend) and on Device::GPU
// (batch B, this backend); the GPU output of sample 0 is compared with the CPU output.
// usage: whisper_encoder [batch [iters [cpu_check 0/1]]]      env AXIOM_CUDA_DEVICE / LOCAL_RANK pick the GPU
#define main whisper_poc_example_main
#include "whisper_poc.cpp"
#undef main

#include "axb.h"
#include "backends/metal/metal_common.hpp"
#include "cuda_plan.hpp"

static void gpu_sync() { backends::metal::MetalExecutionStream::instance().synchronize(); }

static std::map<std::string, Tensor> encoder_state(const WhisperConfig &cfg) {
    const auto d = (size_t)cfg.d_model, ffn = (size_t)cfg.encoder_ffn_dim, mels = (size_t)cfg.n_mels;
    std::map<std::string, Tensor> st;
    st["conv1.weight"] = Tensor::randn({d, mels, 3});
    st["conv1.bias"] = Tensor::zeros({d});
    st["conv2.weight"] = Tensor::randn({d, d, 3});
    st["conv2.bias"] = Tensor::zeros({d});
    st["embed_positions.weight"] = Tensor::randn({(size_t)cfg.max_source_positions, d});
    st["layer_norm.weight"] = Tensor::ones({d});
    st["layer_norm.bias"] = Tensor::zeros({d});
    for (int i = 0; i < cfg.encoder_layers; ++i) {
        const std::string p = "layers." + std::to_string(i) + ".";
        for (auto proj : {"q_proj", "k_proj", "v_proj", "out_proj"}) {
            st[p + "self_attn." + proj + ".weight"] = Tensor::randn({d, d}) * 0.02f;
            st[p + "self_attn." + proj + ".bias"] = Tensor::zeros({d});
        }
        st[p + "self_attn_layer_norm.weight"] = Tensor::ones({d});
        st[p + "self_attn_layer_norm.bias"] = Tensor::zeros({d});
        st[p + "fc1.weight"] = Tensor::randn({ffn, d}) * 0.02f;
        st[p + "fc1.bias"] = Tensor::zeros({ffn});
        st[p + "fc2.weight"] = Tensor::randn({d, ffn}) * 0.02f;
        st[p + "fc2.bias"] = Tensor::zeros({d});
        st[p + "final_layer_norm.weight"] = Tensor::ones({d});
        st[p + "final_layer_norm.bias"] = Tensor::zeros({d});
    }
    return st;
}

int main(int argc, char **argv) {
    const size_t B = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 8;
    const int iters = argc > 2 ? std::atoi(argv[2]) : 5;
    const bool cpu_check = argc > 3 ? std::atoi(argv[3]) != 0 : true;
    using Clock = std::chrono::steady_clock;
    ops::OperationRegistry::initialize_builtin_operations();
    manual_seed(20240917);
    WhisperConfig cfg;
    auto state = encoder_state(cfg);
    auto mel = Tensor::randn({B, (size_t)cfg.n_mels, 3000});

    double cpu_ms = 0, max_abs = 0, max_ref = 0;
    bool ok = true;
    Tensor cpu_out;
    if (cpu_check) {
        WhisperEncoder enc_cpu(cfg);
        enc_cpu.load_state_dict(state, "", false);
        Tensor mel0 = mel.slice({Slice(0, 1)}).ascontiguousarray();
        auto t0 = Clock::now();
        cpu_out = enc_cpu.forward(mel0);
        cpu_out.sync();
        cpu_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
    }

    WhisperEncoder enc(cfg);
    enc.load_state_dict(state, "", false);
    enc.to(Device::GPU);
    Tensor mel_g = mel.gpu();
    Tensor out;
    for (int i = 0; i < 2; ++i) { // warm-up (kernel specialisation, allocator pool)
        out = enc.forward(mel_g);
        out.sync();
    }
    gpu_sync();
    const uint64_t l0 = axb_launch_count();
    // three clocks: wall time of the whole loop (host + device, what a caller sees), the HOST time to build, lower and enqueue
    // the forwards (the loop body never waits for the device), and the DEVICE time between two events on the backend stream
    axb_event_t ev0 = nullptr, ev1 = nullptr;
    axb_event_create(&ev0);
    axb_event_create(&ev1);
    const auto plan0 = backends::cuda::plan::stats();
    axb_event_record(ev0, nullptr);
    auto t0 = Clock::now();
    for (int i = 0; i < iters; ++i) {
        out = enc.forward(mel_g);
        out.sync();
    }
    const double host_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count() / iters;
    axb_event_record(ev1, nullptr);
    gpu_sync();
    const double gpu_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count() / iters;
    float dev_total = 0.0f;
    axb_event_elapsed_ms(ev0, ev1, &dev_total);
    const double device_ms = dev_total / iters;
    const auto plan1 = backends::cuda::plan::stats();
    const uint64_t launches = (axb_launch_count() - l0) / (uint64_t)iters;

    if (cpu_check) {
        Tensor g0 = out.slice({Slice(0, 1)}).cpu().ascontiguousarray();
        Tensor c0 = cpu_out.ascontiguousarray();
        const float *g = g0.typed_data<float>(), *c = c0.typed_data<float>();
        for (size_t i = 0; i < c0.size(); ++i) {
            max_abs = std::max(max_abs, (double)std::fabs(g[i] - c[i]));
            max_ref = std::max(max_ref, (double)std::fabs(c[i]));
        }
        // the output is a LayerNorm of O(1) values after 4 residual blocks of fp32 GEMMs with different summation orders
        ok = g0.shape() == c0.shape() && max_abs <= 2e-3 * std::max(1.0, max_ref) && out.device() == Device::GPU;
    }
    const double gflop_per_sample = 36.9;
    std::printf("{\"workload\": \"whisper_poc encoder forward, random-init, mel {%zu,80,3000}\", \"batch\": %zu, \"gpu_ms\": %.3f,"
                " \"samples_per_s\": %.2f, \"gpu_TFLOPs\": %.2f, \"gpu_launches_per_forward\": %llu, \"cpu_ms_one_sample\": %.1f,"
                " \"host_enqueue_ms\": %.3f, \"device_ms\": %.3f, \"host_over_device\": %.3f, \"plan_cache_hits\": %llu,"
                " \"plan_cache_misses\": %llu,"
                " \"parity_checked\": %s, \"max_abs_diff\": %.3e, \"max_abs_ref\": %.3e, \"parity_ok\": %s}\n",
                B, B, gpu_ms, B / gpu_ms * 1e3, gflop_per_sample * B / gpu_ms, (unsigned long long)launches, cpu_ms, host_ms, device_ms,
                device_ms > 0 ? host_ms / device_ms : 0.0, (unsigned long long)(plan1.hits - plan0.hits),
                (unsigned long long)(plan1.misses - plan0.misses),
                cpu_check ? "true" : "false", max_abs, max_ref, ok ? "true" : "false");
    return ok ? 0 : 1;
}
