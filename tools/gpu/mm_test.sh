timeout 900 python -m pytest tests/test_matmul_gpu.py -m gpu -q 2>&1 | grep -E "FAILED|passed|failed|Error|assert" | head -20
