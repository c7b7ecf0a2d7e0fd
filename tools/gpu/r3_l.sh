AXB_CLUSTER_NO_PIECES=1 timeout 200 python tools/rows_probe.py 8192 32768 | tail -1
timeout 200 python tools/rows_probe.py 8192 32768 | tail -1
AXB_CLUSTER_NO_PIECES=1 timeout 200 python tools/rows_probe.py 8192 32768 | tail -1
