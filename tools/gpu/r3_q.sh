timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k 'masked_inf_nan or two_outer' 2>&1 | tail -15
