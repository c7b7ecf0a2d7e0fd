set -x
timeout 900 python -m pytest tests/test_transfers_gpu.py -q -x --timeout 600 -s 2>&1 | tail -8
AXB_NO_STAGED_COPY=1 timeout 900 python -m pytest tests/test_transfers_gpu.py -q -x --timeout 600 -s -k faster 2>&1 | grep -E "pageable|passed|failed"
