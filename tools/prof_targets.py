"""Runs each hot kernel a few times at its BASELINE shape (for `ncu -k regex:<name>` captures; never a bench number)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
which = sys.argv[1:] or ["c2", "sum", "argmax", "softmax", "layer_norm", "bmm"]
reps = int(os.environ.get("REPS", "3"))
by = ctypes.byref
if "c2" in which:
    import __graft_entry__ as g
    a = dev.randn((4096, 1), 1); b = dev.randn((1, 16384), 2); c = dev.randn((4096, 16384), 3)
    ins = [a, dev.DeviceArray.from_numpy(np.array([1.0], np.float32)), b, c, dev.DeviceArray.from_numpy(np.array([0], np.int32)), dev.DeviceArray.from_numpy(np.array([0.0], np.float32))]
    out = dev.DeviceArray.empty((4096, 16384), capi.F32)
    for _ in range(reps): dev.elementwise(g._config2_program(capi), ins, out=out)
rows, n = 8192, 32768
for dt in ("f32", "bf16"):
    if not any(w in which for w in ("sum", "argmax", "softmax", "layer_norm")): break
    if dt == "bf16" and "bf16" not in which: continue
    x = dev.randn((rows, n), 7)
    if dt == "bf16": x = dev.cast(x, capi.BF16)
    w, b = dev.randn((n,), 8, 1.0, 0.1), dev.randn((n,), 9, 0.0, 0.1)
    for _ in range(reps):
        if "sum" in which: dev.reduce("sum", x, [1])
        if "argmax" in which: dev.argreduce("argmax", x, 1)
        if "softmax" in which: dev.softmax(x, -1)
        if "layer_norm" in which: dev.layer_norm(x, w, b, -1, 1e-5)
    del x
if "bmm" in which:
    A = dev.cast(dev.randn((8, 4096, 4096), 11), capi.BF16); B = dev.cast(dev.randn((8, 4096, 4096), 12, 0.0, 1 / 64), capi.BF16)
    for _ in range(reps): dev.matmul(A, B)
if "mm32" in which:
    A = dev.randn((4096, 4096), 31); B = dev.randn((4096, 4096), 32, 0.0, 1 / 64)
    for _ in range(reps): dev.matmul(A, B)
dev.synchronize()
print("ok")
