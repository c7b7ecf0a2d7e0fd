"""Small invocations of the round-1 kernels for compute-sanitizer (memcheck / racecheck / synccheck); not a test of values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
capi.check(capi.lib().axb_set_device(0))
rng = np.random.default_rng(0)
f = lambda *s: dev.DeviceArray.from_numpy(rng.standard_normal(s).astype(np.float32))
x = f(37, 5000); dev.softmax(x, -1); dev.softmax(x, -1, True)                      # wide, ragged lane map
x = f(150, 8192); dev.softmax(x, -1); dev.layer_norm(x, None, None, -1, 1e-5)       # wide J=2
x = f(600, 8192); w = f(8192); b = f(8192); dev.layer_norm(x, w, b, -1, 1e-5); dev.rms_norm(x, w, -1, 1e-5)  # cluster
xb = dev.cast(f(601, 16384), capi.BF16); w = f(16384); b = f(16384); dev.layer_norm(xb, w, b, -1, 1e-5)     # cluster R=2, ragged
x = f(64, 4096); dev.softmax(x, -1); x = f(40, 1500); dev.softmax(x, -1); dev.layer_norm(x, f(1500), f(1500), -1, 1e-5)
x = f(33, 70001); dev.argreduce("argmax", x, 1); dev.reduce("sum", x, [1]); dev.reduce("max", x, [0])
a = f(512, 320); bt = f(384, 320); dev.matmul(a, bt, transpose_b=True, bias=f(384), activation="gelu")   # fp32 split GEMM
ab = dev.cast(f(300, 200), capi.BF16); bb = dev.cast(f(200, 260), capi.BF16); dev.matmul(ab, bb)
a3 = f(3, 200, 96); b3 = f(3, 96, 300); dev.matmul(a3, b3)
dev.synchronize()
print("sanitize targets ok")
