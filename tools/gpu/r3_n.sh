timeout 200 python tools/softmax_probe.py 8192 32768 | tail -1
timeout 200 python tools/softmax_probe.py 8192 32768 | tail -1
timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k 'softmax' 2>&1 | tail -1
