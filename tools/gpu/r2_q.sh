set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
cd axiom_b200/host/_build
timeout 300 ./whisper_encoder 1 30 0 2>&1 | tail -1
AXIOM_CUDA_NO_PLAN_CACHE=1 timeout 300 ./whisper_encoder 1 30 0 2>&1 | tail -1
timeout 300 ./whisper_encoder 8 10 0 2>&1 | tail -1
timeout 300 ./axiom_chains 4096 16384 20 2>&1 | tail -1 | cut -c1-400
