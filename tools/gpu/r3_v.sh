timeout 600 python -m pytest tests/test_spatial_gpu.py -m gpu -x -q -k "grouped or bit_exact" 2>&1 | tail -12
