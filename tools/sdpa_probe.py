"""Device time of axb_sdpa at the whisper shape: single flash kernel vs the three-kernel path, fp32 / bf16.
usage: python tools/sdpa_probe.py [B H L D]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev

L = capi.lib()
capi.check(L.axb_set_device(0))
B, H, Ln, D = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (8, 6, 1500, 64)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    capi.check(L.axb_event_create(ctypes.byref(e0))); capi.check(L.axb_event_create(ctypes.byref(e1)))
    capi.check(L.axb_event_record(e0, None))
    for _ in range(iters):
        fn()
    capi.check(L.axb_event_record(e1, None)); capi.check(L.axb_event_sync(e1))
    ms = ctypes.c_float(); capi.check(L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms)))
    return ms.value / iters


out = {"shape": [B, H, Ln, D]}
flops = 4.0 * B * H * Ln * Ln * D
for dt in ("f32", "bf16"):
    q, k, v = (dev.randn((B, H, Ln, D), 40 + i) for i in range(3))
    code = capi.F32
    if dt == "bf16":
        q, k, v = (dev.cast(t, capi.BF16) for t in (q, k, v))
        code = capi.BF16
    o = dev.DeviceArray.empty((B, H, Ln, D), code)
    qd, kd, vd, od = q.desc(), k.desc(), v.desc(), o.desc()
    by = ctypes.byref
    for causal in (0, 1):
        fn = lambda: capi.check(L.axb_sdpa(by(qd), by(kd), by(vd), None, ctypes.c_float(-1.0), causal, by(od), None))
        for path in ("flash", "three_kernel"):
            if path == "three_kernel":
                if dt != "f32":
                    continue
                os.environ["AXB_SDPA_NO_FLASH"] = "1"
            try:
                ms = timeit(fn)
            finally:
                os.environ.pop("AXB_SDPA_NO_FLASH", None)
            eff = flops * (0.5 if causal else 1.0)
            out[f"{dt}_{'causal' if causal else 'full'}_{path}"] = {"ms": round(ms, 4), "TFLOPs_useful": round(eff / ms / 1e9, 1),
                                                                  "bf16_MMA_TFLOPs": round(eff * (6 if dt == "f32" else 1) / ms / 1e9, 1)}
print(json.dumps(out))
