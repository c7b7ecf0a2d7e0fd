timeout 1500 python -m pytest tests/test_elementwise_gpu.py -m gpu -q -k exhaustive 2>&1 | tail -8
