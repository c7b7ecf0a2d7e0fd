timeout 1500 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q 2>&1 | tail -15
