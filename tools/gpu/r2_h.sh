set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reduce_gpu.py -m gpu -q -x --timeout 300 -k "softmax or norm" 2>&1 | tail -4
timeout 300 python tools/rows_probe.py 8192 32768 | tee gpurun_out/r2_rows_cluster.json
AXB_NO_CLUSTER_SOFTMAX=1 timeout 300 python tools/rows_probe.py 8192 32768 | tee gpurun_out/r2_rows_wide.json
timeout 300 python tools/rows_probe.py 4096 32768 | tee gpurun_out/r2_rows_cluster_4096.json
timeout 300 python tools/rows_probe.py 8192 16384 | tee gpurun_out/r2_rows_cluster_n16384.json
AXB_NO_CLUSTER_SOFTMAX=1 timeout 300 python tools/rows_probe.py 8192 16384 | tee gpurun_out/r2_rows_wide_n16384.json
