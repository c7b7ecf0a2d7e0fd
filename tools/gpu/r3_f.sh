set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k "softmax or norm" 2>&1 | tail -3
o=gpurun_out/r3_rows_probe_f.json
timeout 200 python tools/rows_probe.py 8192 32768 | tail -1 | tee -a $o
AXB_CLUSTER_NO_PIECES=1 timeout 200 python tools/rows_probe.py 8192 32768 | tail -1 | tee -a $o
