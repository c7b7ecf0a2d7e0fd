mkdir -p gpurun_out
cd axiom_b200/host/_build
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /root/repo/gpurun_out/r3_whisper_launches_b8_d.csv ./whisper_encoder 8 1 0 > /root/repo/gpurun_out/whisper_ncu.log 2>&1
wc -l /root/repo/gpurun_out/r3_whisper_launches_b8_d.csv
timeout 300 ./whisper_encoder 1 10 0 2>&1 | tail -1 | cut -c1-400
