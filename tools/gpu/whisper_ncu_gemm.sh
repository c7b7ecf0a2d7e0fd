cd axiom_b200/host/_build
ncu --set full --clock-control none -k regex:gemm_tc --launch-skip 90 --launch-count 6 -o /root/repo/gpurun_out/r1s6_whisper_gemm ./whisper_encoder 8 1 0 > /root/repo/gpurun_out/whisper_ncu2.log 2>&1
tail -2 /root/repo/gpurun_out/whisper_ncu2.log | cut -c1-200
