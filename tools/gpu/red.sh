timeout 600 python -m pytest tests/test_reduce_gpu.py -m gpu -q -k "fused_chain" 2>&1 | tail -8
cd axiom_b200/host/_build; for t in test_fusion test_fusion_extended test_lazy_evaluation test_graph_compiler test_cpu_gpu_parity; do timeout 300 ./reftest_$t 2>&1 | grep -E "^\[  FAILED|MINIGTEST_SUMMARY"; done; cd ../../..
python - <<'PY'
import ctypes, time, numpy as np, os
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
F, V = capi.F32, capi.SRC_VALUE
n = 1 << 26
a, b, c = dev.randn((n,), 1), dev.randn((n,), 2), dev.randn((n,), 3)
prog = [capi.make_instr(capi.OP["add"], F, 0, 1), capi.make_instr(capi.OP["mul"], F, V(0), 2)]
e0, e1 = ctypes.c_void_p(), ctypes.c_void_p(); L.axb_event_create(ctypes.byref(e0)); L.axb_event_create(ctypes.byref(e1))
for _ in range(3): dev.elementwise_reduce(prog, [a, b, c], "sum")
L.axb_event_record(e0, None)
for _ in range(20): dev.elementwise_reduce(prog, [a, b, c], "sum")
L.axb_event_record(e1, None); L.axb_event_sync(e1)
ms = ctypes.c_float(); L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms))
print("sum((a+b)*c) 3 x 256 MiB: %.3f ms/call, %.0f GB/s algorithmic" % (ms.value / 20, 3 * n * 4 / (ms.value / 20) / 1e6))
PY
AXB_EW_REDUCE_TWO_PASS=1 python - <<'PY'
import ctypes, numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
F, V = capi.F32, capi.SRC_VALUE
n = 1 << 26
a, b, c = dev.randn((n,), 1), dev.randn((n,), 2), dev.randn((n,), 3)
prog = [capi.make_instr(capi.OP["add"], F, 0, 1), capi.make_instr(capi.OP["mul"], F, V(0), 2)]
e0, e1 = ctypes.c_void_p(), ctypes.c_void_p(); L.axb_event_create(ctypes.byref(e0)); L.axb_event_create(ctypes.byref(e1))
for _ in range(3): dev.elementwise_reduce(prog, [a, b, c], "sum")
L.axb_event_record(e0, None)
for _ in range(20): dev.elementwise_reduce(prog, [a, b, c], "sum")
L.axb_event_record(e1, None); L.axb_event_sync(e1)
ms = ctypes.c_float(); L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms))
print("two-pass (chain materialised): %.3f ms/call" % (ms.value / 20))
PY
