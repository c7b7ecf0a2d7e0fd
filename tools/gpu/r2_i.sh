set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_elementwise_gpu.py tests/test_reduce_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 300 axiom_b200/host/_build/axiom_chains 4096 16384 20 2>&1 | tail -1
timeout 300 axiom_b200/host/_build/whisper_encoder 1 20 0 2>&1 | tail -1
AXB_NO_PLAN_CACHE=1 timeout 300 axiom_b200/host/_build/whisper_encoder 1 20 0 2>&1 | tail -1
timeout 300 axiom_b200/host/_build/whisper_encoder 8 5 0 2>&1 | tail -1
