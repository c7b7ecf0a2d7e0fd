timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -q -k sharded 2>&1 | tail -12
