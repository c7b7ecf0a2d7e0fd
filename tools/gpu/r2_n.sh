set -x
timeout 1500 python -m pytest tests/test_spatial_gpu.py tests/test_axiom_cpp.py tests/test_attention_gpu.py -q -x --timeout 900 2>&1 | tail -8
cd axiom_b200/host/_build
timeout 300 ./whisper_encoder 8 10 0 2>&1 | tail -1
timeout 300 ./whisper_encoder 1 20 0 2>&1 | tail -1
timeout 600 ./whisper_encoder 2 1 1 2>&1 | tail -1
