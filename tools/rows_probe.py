"""C3 row kernels at {rows, 32768}: softmax / log_softmax / layer_norm, fp32 and bf16, burst timing (best of 5 x 10 calls),
with the cluster softmax on and off (AXB_NO_CLUSTER_SOFTMAX is read once per process, so the A/B is two invocations)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
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
out = {"rows": rows, "n": n, "cluster_softmax": os.environ.get("AXB_NO_CLUSTER_SOFTMAX") is None}
by = ctypes.byref
for dt in ("f32", "bf16"):
    x = dev.randn((rows, n), 7); isz = 4
    if dt == "bf16": x = dev.cast(x, capi.BF16); isz = 2
    w, b = dev.randn((n,), 8, 1.0, 0.1), dev.randn((n,), 9, 0.0, 0.1)
    o = dev.DeviceArray.empty((rows, n), x.dtype); oa = dev.DeviceArray.empty((rows,), capi.I64)
    xd, od, wd, bd, oad = x.desc(), o.desc(), w.desc(), b.desc(), oa.desc()
    for name, fn, nb in (("softmax", lambda: capi.check(L.axb_softmax(by(xd), -1, 0, by(od), None)), 2*rows*n*isz),
                         ("log_softmax", lambda: capi.check(L.axb_softmax(by(xd), -1, 1, by(od), None)), 2*rows*n*isz),
                         ("layer_norm", lambda: capi.check(L.axb_layer_norm(by(xd), by(wd), by(bd), -1, 1e-5, by(od), None)), 2*rows*n*isz),
                         ("argmax", lambda: capi.check(L.axb_argreduce(capi.OP["argmax"], by(xd), 1, 0, by(oad), None)), rows*n*isz)):
        ms = timeit(fn)
        out[f"{name}_{dt}"] = {"ms": round(ms, 4), "GBps": round(nb/ms/1e6, 1), "frac": round(nb/ms/1e6/PEAK, 3), "path": int(L.axb_rownorm_last_path())}
print(json.dumps(out))
