cd axiom_b200/host/_build
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /root/repo/gpurun_out/whisper_launches.csv ./whisper_encoder 8 1 0 > /root/repo/gpurun_out/whisper_ncu.log 2>&1
tail -2 /root/repo/gpurun_out/whisper_ncu.log
