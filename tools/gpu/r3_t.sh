mkdir -p gpurun_out
timeout 300 python tools/convt_probe.py | tail -1 | tee -a gpurun_out/r3_convt_probe.json
timeout 600 python -m pytest tests/test_spatial_gpu.py -m gpu -x -q -k "transpose" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_convt_launches.csv python tools/convt_probe.py > /dev/null 2>&1
wc -l gpurun_out/r3_convt_launches.csv
