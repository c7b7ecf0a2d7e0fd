set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -4
timeout 300 python tools/sdpa_probe.py 8 6 1500 64 > gpurun_out/r2_sdpa_probe_b8.json 2> gpurun_out/r2_sdpa_probe.err; cat gpurun_out/r2_sdpa_probe_b8.json; tail -3 gpurun_out/r2_sdpa_probe.err
timeout 300 axiom_b200/host/_build/whisper_encoder 8 5 1 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:fa_kernel -s 2 -c 1 -o gpurun_out/r2_fa_prof3 python tools/sdpa_probe.py 8 6 1500 64 > gpurun_out/r2_fa_prof.log 2>&1; ls -la gpurun_out/*.ncu-rep
