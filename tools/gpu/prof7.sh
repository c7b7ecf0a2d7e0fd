timeout 600 ncu --set full --clock-control none -k regex:"axb_ew" -c 16 -o /tmp/r1s7_ew python tools/prof_targets_s7.py > gpurun_out/ncu_ew7.log 2>&1; tail -2 gpurun_out/ncu_ew7.log
python tools/ncu_summary.py /tmp/r1s7_ew.ncu-rep > gpurun_out/ncu_summary_s7_ew.txt 2>&1; wc -l gpurun_out/ncu_summary_s7_ew.txt
