# round 2, first pass: cost of the UNPATCHED literal config-2 call, then the patched build, the new API drivers, the whole
# -m gpu suite, and the bench line with the new extras
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 axiom_b200/host/_build_nopatch/axiom_chains 4096 16384 5 > gpurun_out/r2_chains_unpatched.json 2> gpurun_out/r2_chains_unpatched.err; cat gpurun_out/r2_chains_unpatched.json; tail -3 gpurun_out/r2_chains_unpatched.err
timeout 600 axiom_b200/host/_build/axiom_chains 4096 16384 20 > gpurun_out/r2_chains.json 2> gpurun_out/r2_chains.err; cat gpurun_out/r2_chains.json; tail -3 gpurun_out/r2_chains.err
timeout 600 axiom_b200/host/_build/api_parity > gpurun_out/r2_api_parity.json 2>&1; cat gpurun_out/r2_api_parity.json
timeout 1200 axiom_b200/host/_build/c3_c4_api > gpurun_out/r2_c3_c4_api.json 2> gpurun_out/r2_c3_c4_api.err; cat gpurun_out/r2_c3_c4_api.json; tail -3 gpurun_out/r2_c3_c4_api.err
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -3 gpurun_out/r2_bench_a.err; cut -c1-3000 gpurun_out/r2_bench_a.json
