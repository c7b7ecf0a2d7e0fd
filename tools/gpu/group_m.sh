for g in 2 4 8 16; do echo "GROUP_M=$g"; AXB_GEMM_GROUP_M=$g timeout 200 python tools/perf_probe.py bmm mm:bf16:8192:8192:8192 mm:bf16:16384:16384:4096 2>&1 | cut -c1-160; done
