mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_conv_launches2.csv python tools/perf_probe.py conv:bf16:32,256,56,56:256,256,3,3:1:1 conv:bf16:32,512,14,14:512,512,3,3:1:1 > gpurun_out/conv_ncu.log 2>&1
grep -E "conv_prep|gemm_tc" gpurun_out/r2_conv_launches2.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | awk '{print $(NF-1), $NF, substr($1,1,40)}' | sed -n '1,12p;200,212p'
