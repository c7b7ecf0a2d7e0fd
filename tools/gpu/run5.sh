for st in 1 2 3; do for alt in 0 1; do
echo "stages=$st noalt=$alt"
if [ $alt = 1 ]; then export AXB_WB_NO_ALTERNATE=1; else unset AXB_WB_NO_ALTERNATE; fi
AXB_WIDE_STAGES=$st python tools/perf_probe.py c3:layer_norm:bf16 c3:rms_norm:bf16 c3:softmax:bf16
done; done
