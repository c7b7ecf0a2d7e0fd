timeout 900 python -m pytest tests/test_spatial_gpu.py -m gpu -x -q 2>&1 | tail -25
