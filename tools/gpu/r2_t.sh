set -x
timeout 1200 python -m pytest tests/test_elementwise_gpu.py -q -x --timeout 600 2>&1 | tail -2
timeout 300 python tools/perf_probe.py sigmoid silu 2>&1 | tail -2
cd axiom_b200/host/_build
timeout 300 ./whisper_encoder 8 10 0 2>&1 | tail -1 | cut -c1-330
timeout 600 ./whisper_encoder 2 1 1 2>&1 | tail -1 | cut -c300-600
timeout 600 python /root/repo/tests/../tools/perf_probe.py relu 2>&1 | tail -1
