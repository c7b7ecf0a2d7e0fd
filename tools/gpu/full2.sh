set -x
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
