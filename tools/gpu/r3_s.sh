mkdir -p gpurun_out
timeout 300 python tools/convt_probe.py | tail -1 | tee -a gpurun_out/r3_convt_probe.json
AXB_CONVT_NO_TC=1 timeout 600 python tools/convt_probe.py | tail -1 | tee -a gpurun_out/r3_convt_probe.json
