"""softmax / log_softmax at {rows, n}, fp32 and bf16: burst timing plus a hash of the whole result, so that A/B switches
(AXB_WIDE_PIECES, AXB_WIDE_STAGES) can be compared for speed AND for bit-identical output in separate invocations."""
import ctypes, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
dts = sys.argv[3].split(",") if len(sys.argv) > 3 else ["f32", "bf16"]
PEAK = 6551.4
def timeit(fn, k=10, reps=5):
    for _ in range(3): fn()
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    capi.check(L.axb_event_create(ctypes.byref(e0))); capi.check(L.axb_event_create(ctypes.byref(e1)))
    best = 1e9
    for _ in range(reps):
        capi.check(L.axb_event_record(e0, None))
        for _ in range(k): fn()
        capi.check(L.axb_event_record(e1, None)); capi.check(L.axb_event_sync(e1))
        ms = ctypes.c_float(); capi.check(L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms))); best = min(best, ms.value / k)
    return best
out = {"rows": rows, "n": n, "pieces": os.environ.get("AXB_WIDE_PIECES"), "stages": os.environ.get("AXB_WIDE_STAGES")}
by = ctypes.byref
for dt in dts:
    x = dev.randn((rows, n), 7); isz = 4
    if dt == "bf16": x = dev.cast(x, capi.BF16); isz = 2
    o = dev.DeviceArray.empty((rows, n), x.dtype)
    xd, od = x.desc(), o.desc()
    for name, lg in (("softmax", 0), ("log_softmax", 1)):
        fn = lambda: capi.check(L.axb_softmax(by(xd), -1, lg, by(od), None))
        ms = timeit(fn)
        h = hashlib.sha1(o.numpy().tobytes()).hexdigest()[:12]
        out[f"{name}_{dt}"] = {"ms": round(ms, 4), "GBps": round(2*rows*n*isz/ms/1e6, 1), "frac": round(2*rows*n*isz/ms/1e6/PEAK, 3),
                               "path": int(L.axb_rownorm_last_path()), "sha": h}
print(json.dumps(out))
