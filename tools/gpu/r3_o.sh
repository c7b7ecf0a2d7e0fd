for i in 1 2; do timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k 'deterministic' 2>&1 | tail -3; done
