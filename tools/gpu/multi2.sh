set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err | cut -c1-300; cat gpurun_out/bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_ref_n2.json 2>> gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json | cut -c1-400
