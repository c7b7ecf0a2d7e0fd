set -x
mkdir -p gpurun_out
timeout 300 python tools/sdpa_probe.py 8 6 1500 64 > gpurun_out/r2_sdpa_probe_b8.json 2> gpurun_out/r2_sdpa_probe.err; cat gpurun_out/r2_sdpa_probe_b8.json; tail -3 gpurun_out/r2_sdpa_probe.err
timeout 300 python tools/sdpa_probe.py 32 6 1500 64 > gpurun_out/r2_sdpa_probe_b32.json 2>> gpurun_out/r2_sdpa_probe.err; cat gpurun_out/r2_sdpa_probe_b32.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_whisper_b8.csv axiom_b200/host/_build/whisper_encoder 8 3 0 > gpurun_out/r2_whisper_under_ncu.log 2>&1; wc -l gpurun_out/r2_launches_whisper_b8.csv
timeout 300 axiom_b200/host/_build/whisper_encoder 32 3 0 2>&1 | tail -1
timeout 300 axiom_b200/host/_build/whisper_encoder 1 10 0 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:fa_kernel -s 2 -c 1 -o gpurun_out/r2_fa_prof python tools/sdpa_probe.py 8 6 1500 64 > gpurun_out/r2_fa_prof.log 2>&1; ls -la gpurun_out/*.ncu-rep
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
