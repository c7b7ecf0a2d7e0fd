timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_matmul_gpu.py -m gpu -x -q 2>&1 | tail -3
cd axiom_b200/host/_build; timeout 300 ./whisper_encoder 8 5 1 | tail -1; timeout 300 ./whisper_encoder 32 3 0 | tail -1
