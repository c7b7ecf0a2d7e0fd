set -x
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -k regex:rowwide -s 6 -c 1"
run() { # name, dtype
  timeout 300 $N -o /tmp/$1 python tools/softmax_probe.py 8192 32768 $2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/r3_ncu_$1.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/r3_src_$1.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r3_raw_$1.csv 2>/dev/null
}
run sm_bf16_c1 bf16
AXB_WIDE_PIECES=2 AXB_WIDE_STAGES=4 run sm_bf16_c2 bf16
run sm_f32_c1 f32
AXB_WIDE_PIECES=4 AXB_WIDE_STAGES=4 run sm_f32_c4 f32
ls -la gpurun_out/
