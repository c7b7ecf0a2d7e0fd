set -x
mkdir -p gpurun_out
P="timeout 200 python tools/softmax_probe.py"
o=gpurun_out/r3_softmax_probe.json
$P 8192 32768 | tail -1 | tee -a $o
for pc in 2 4 8; do AXB_WIDE_PIECES=$pc $P 8192 32768 f32 | tail -1 | tee -a $o; done
AXB_WIDE_PIECES=2 AXB_WIDE_STAGES=2 $P 8192 32768 f32 | tail -1 | tee -a $o
AXB_WIDE_PIECES=8 AXB_WIDE_STAGES=8 $P 8192 32768 f32 | tail -1 | tee -a $o
AXB_WIDE_PIECES=8 AXB_WIDE_STAGES=10 $P 8192 32768 f32 | tail -1 | tee -a $o
for st in 2 3; do AXB_WIDE_STAGES=$st $P 8192 32768 bf16 | tail -1 | tee -a $o; done
for n in 32512 31744 28672 30720; do $P 8192 $n f32 | tail -1 | tee -a $o; done
AXB_WIDE_STAGES=2 $P 16384 16384 f32 | tail -1 | tee -a $o
AXB_WIDE_STAGES=3 $P 16384 16384 f32 | tail -1 | tee -a $o
