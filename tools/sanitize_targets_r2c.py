"""Small invocations of the kernels touched in the last session of round 2 for compute-sanitizer (memcheck / racecheck /
synccheck): TMA-staged row kernels (four-piece fetch of 128 KiB fp32 rows, deferred refill, bf16 rows, ragged widths), the
SM-pair cluster layer_norm / rms_norm, the fp32 GEMM with the MUFU GELU epilogue, fp32 erf / gelu chains. Not a test of values."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
rng = np.random.default_rng(0)
paths = set()
f = lambda *s: dev.DeviceArray.from_numpy(rng.standard_normal(s).astype(np.float32))
for shp in ((300, 32768), (3, 50, 32768), (310, 20000), (297, 8192)):
    x = f(*shp)
    for lg in (False, True):
        dev.softmax(x, -1, log=lg); paths.add((shp, 'softmax', int(L.axb_rownorm_last_path())))
        dev.softmax(dev.cast(x, capi.BF16), -1, log=lg)
    dev.layer_norm(x, None, None, -1, 1e-5); dev.rms_norm(x, None, -1, 1e-5)
for shp in ((700, 32768), (601, 16384), (593, 8320)):
    x, w, b = f(*shp), f(shp[-1]), f(shp[-1])
    dev.layer_norm(x, w, b, -1, 1e-5); paths.add((shp, 'layer_norm', int(L.axb_rownorm_last_path())))
    dev.rms_norm(x, w, -1, 1e-5)
    dev.layer_norm(dev.cast(x, capi.BF16), w, b, -1, 1e-5)
a, w, bias = f(1000, 384), f(1536, 384), f(1536)
dev.matmul(a, w, transpose_b=True, bias=bias, activation="gelu"); paths.add(("mm", "gelu", int(L.axb_matmul_last_path())))
xe = f((1 << 20) + 5); dev.unary("erf", xe); dev.unary("gelu", xe)
dev.synchronize()
print("sanitize targets ok", sorted(paths, key=str))
