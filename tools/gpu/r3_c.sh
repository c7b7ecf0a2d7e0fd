set -x
mkdir -p gpurun_out
P="timeout 200 python tools/softmax_probe.py"
o=gpurun_out/r3_softmax_probe_c.json
for st in 3 4 5 6; do AXB_WIDE_PIECES=4 AXB_WIDE_STAGES=$st $P 8192 32768 f32 | tail -1 | tee -a $o; done
for pc in 2 4; do for st in 2 3 4 6 8 12; do AXB_WIDE_PIECES=$pc AXB_WIDE_STAGES=$st $P 8192 32768 bf16 | tail -1 | tee -a $o; done; done
