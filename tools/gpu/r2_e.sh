set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -8
timeout 300 python tools/sdpa_probe.py 8 6 1500 64 > gpurun_out/r2_sdpa_probe_b8.json 2> gpurun_out/r2_sdpa_probe.err; cat gpurun_out/r2_sdpa_probe_b8.json; tail -3 gpurun_out/r2_sdpa_probe.err
timeout 300 python tools/sdpa_probe.py 32 6 1500 64 > gpurun_out/r2_sdpa_probe_b32.json 2>> gpurun_out/r2_sdpa_probe.err; cat gpurun_out/r2_sdpa_probe_b32.json
timeout 300 axiom_b200/host/_build/whisper_encoder 8 5 1 2>&1 | tail -1
timeout 300 axiom_b200/host/_build/reftest_test_flash_attention 2>&1 | tail -2
