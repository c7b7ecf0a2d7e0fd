python tools/perf_probe.py mm:f32:12000:384:384 mm:f32:12000:384:384:1 mm:f32:12000:1536:384:1:gelu mm:f32:12000:384:1536:1 mm:bf16:12000:384:384 mm:bf16:12000:384:384:1 mm:bf16:12000:1536:384:1:gelu mm:bf16:12032:512:2304 mm:f32:4096:4096:4096 mm:bf16:4096:4096:4096
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 3 --launch-count 1 -o gpurun_out/r1s6_gemm_probe python tools/perf_probe.py mm:f32:12000:384:384:1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
