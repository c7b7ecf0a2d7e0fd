timeout 300 python tools/ln_race_probe.py 2>&1 | tail -12 | tr '\n' ' '; echo
for i in 1 2 3; do echo "run $i: $(timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k 'softmax or norm' 2>&1 | tail -1)"; done
timeout 200 python tools/rows_probe.py 8192 32768 | tail -1
timeout 200 python tools/softmax_probe.py 8192 32768 | tail -1
