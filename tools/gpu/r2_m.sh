set -x
timeout 1200 python -m pytest tests/test_spatial_gpu.py -q -x --timeout 600 -k "implicit" 2>&1 | tail -25
