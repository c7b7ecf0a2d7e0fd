set -x
mkdir -p gpurun_out
timeout 600 axiom_b200/host/_build/axiom_chains 4096 16384 20 > gpurun_out/r2_chains.json 2> gpurun_out/r2_chains.err; cat gpurun_out/r2_chains.json; tail -3 gpurun_out/r2_chains.err
timeout 600 axiom_b200/host/_build/api_parity > gpurun_out/r2_api_parity.json 2>&1; cat gpurun_out/r2_api_parity.json
timeout 1200 axiom_b200/host/_build/c3_c4_api 8192 32768 64 4096 5 > gpurun_out/r2_c3_c4_api.json 2> gpurun_out/r2_c3_c4_api.err; cut -c1-400 gpurun_out/r2_c3_c4_api.json; tail -3 gpurun_out/r2_c3_c4_api.err
timeout 300 python tools/power_probe.py 4 ours,cublas > gpurun_out/r2_power_probe_a.json 2> gpurun_out/r2_power_probe_a.err; cut -c1-1500 gpurun_out/r2_power_probe_a.json; tail -3 gpurun_out/r2_power_probe_a.err
timeout 300 python tools/power_probe.py 4 cublas,ours > gpurun_out/r2_power_probe_b.json 2> gpurun_out/r2_power_probe_b.err; cut -c1-600 gpurun_out/r2_power_probe_b.json
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15
