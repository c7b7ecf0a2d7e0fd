"""Transposed convolution: tensor-core path (one implicit-GEMM convolution per output phase) against the gather kernel
(AXB_CONVT_NO_TC=1, read once per process, so the A/B is two invocations). Burst timing, best of 5 x 5 calls."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
def timeit(fn, k=5, reps=5):
    for _ in range(2): fn()
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    capi.check(L.axb_event_create(ctypes.byref(e0))); capi.check(L.axb_event_create(ctypes.byref(e1)))
    best = 1e9
    for _ in range(reps):
        capi.check(L.axb_event_record(e0, None))
        for _ in range(k): fn()
        capi.check(L.axb_event_record(e1, None)); capi.check(L.axb_event_sync(e1))
        ms = ctypes.c_float(); capi.check(L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms))); best = min(best, ms.value / k)
    return best
cases = [("2d_256to128_32x32_k4s2p1", (8, 256, 32, 32), (256, 128, 4, 4), dict(stride=2, padding=1)),
         ("2d_128to64_64x64_k3s2p1op1", (8, 128, 64, 64), (128, 64, 3, 3), dict(stride=2, padding=1, output_padding=1)),
         ("1d_512to256_L2000_k16s8p4", (8, 512, 2000), (512, 256, 16), dict(stride=8, padding=4))]
out = {"tc": os.environ.get("AXB_CONVT_NO_TC") is None}
for name, xs, ws, kw in cases:
    for dt in ("f32", "bf16"):
        x, w = dev.randn(xs, 1), dev.randn(ws, 2, 0.0, 0.05)
        if dt == "bf16": x, w = dev.cast(x, capi.BF16), dev.cast(w, capi.BF16)
        fn = lambda: dev.conv_transpose(x, w, None, **kw)
        ms = timeit(fn)
        flops = 2.0 * np.prod(xs) * ws[1] * np.prod(ws[2:])
        out[f"{name}_{dt}"] = {"ms": round(ms, 4), "TFLOPs": round(flops / ms / 1e9, 1), "path": int(L.axb_matmul_last_path())}
print(json.dumps(out))
