set -x
timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -q -k sharded 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err | cut -c1-300; cat gpurun_out/bench_n2.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']); print(json.dumps(d['extras']['c3_full_sum_row_sharded']))"
