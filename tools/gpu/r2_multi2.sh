set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 2>&1 | tail -6
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -5 gpurun_out/r2_bench_n2.err; tail -c 3000 gpurun_out/r2_bench_n2.json
