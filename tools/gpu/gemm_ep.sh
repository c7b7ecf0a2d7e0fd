timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/perf_probe.py mm:f32:12000:384:384:1 mm:f32:12000:1536:384:1:gelu mm:bf16:12000:384:384:1 mm:f32:4096:4096:4096 mm:bf16:4096:4096:4096 bmm
cd axiom_b200/host/_build; timeout 300 ./whisper_encoder 8 5 1 | tail -1
