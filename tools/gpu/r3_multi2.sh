set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r3_bench_n2.json 2> gpurun_out/r3_bench_n2.err; tail -3 gpurun_out/r3_bench_n2.err | cut -c1-300; wc -c gpurun_out/r3_bench_n2.json
