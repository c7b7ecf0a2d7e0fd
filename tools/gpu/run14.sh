timeout 900 python -m pytest tests/test_elementwise_gpu.py -m gpu -x -q -k "host_entry or config2" 2>&1 | tail -12
timeout 600 python bench.py --no-extras --steps 30 > gpurun_out/b14.json 2> gpurun_out/b14.err; tail -5 gpurun_out/b14.err; cut -c1-300 gpurun_out/b14.json; grep -o '"e2e": {[^}]*}' gpurun_out/b14.json
