set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reduce_gpu.py -m gpu -x -q -k "softmax or norm" 2>&1 | tail -3
P="timeout 200 python tools/softmax_probe.py"
o=gpurun_out/r3_softmax_probe_e.json
$P 8192 32768 | tail -1 | tee -a $o
AXB_WIDE_OUT_BUFS=1 $P 8192 32768 bf16 | tail -1 | tee -a $o
AXB_WIDE_STAGES=2 AXB_WIDE_OUT_BUFS=1 $P 8192 32768 bf16 | tail -1 | tee -a $o
AXB_WIDE_NO_TS=1 $P 8192 32768 bf16 | tail -1 | tee -a $o
$P 16384 16384 bf16 | tail -1 | tee -a $o
AXB_WIDE_NO_TS=1 $P 16384 16384 bf16 | tail -1 | tee -a $o
$P 8192 24576 bf16 | tail -1 | tee -a $o
AXB_WIDE_NO_TS=1 $P 8192 24576 bf16 | tail -1 | tee -a $o
