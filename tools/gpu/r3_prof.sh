set -x
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:gemm_tc_kernel -s 8 -c 1 -o /tmp/p_mlp python tools/perf_probe.py mm:f32:12000:1536:384:1:gelu > /dev/null 2>&1
timeout 300 $N -k regex:rowwide -s 6 -c 1 -o /tmp/p_sm python tools/softmax_probe.py 8192 32768 f32 > /dev/null 2>&1
timeout 300 $N -k regex:gemm_tc_kernel -s 6 -c 1 -o /tmp/p_convt python tools/convt_probe.py > /dev/null 2>&1
for f in p_mlp p_sm p_convt; do python tools/ncu_summary.py /tmp/$f.ncu-rep; done > gpurun_out/r3_ncu_final.txt 2>&1
wc -l gpurun_out/r3_ncu_final.txt
