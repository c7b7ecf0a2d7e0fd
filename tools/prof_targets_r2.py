"""ncu targets of round 2: fp32 gelu / exp chains, the half-precision table kernel, the flash-attention kernel at the whisper
shape (fp32 three-piece and bf16), and the fp32 Linear GEMM of the whisper MLP (12000 x 384) . (384 x 1536) + bias + gelu."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
by = ctypes.byref
x = dev.randn((1 << 28,), 5)
for op in ("gelu", "exp"):                      # axb_ew, fp32
    for _ in range(2): dev.unary(op, x)
xb = dev.cast(dev.randn((1 << 28,), 6), capi.BF16)
for _ in range(2): dev.unary("gelu", xb)        # lut_kernel
del x, xb
for dt in (capi.F32, capi.BF16):                # fa_kernel
    q, k, v = (dev.randn((8, 6, 1500, 64), 40 + i) for i in range(3))
    if dt == capi.BF16:
        q, k, v = (dev.cast(t, dt) for t in (q, k, v))
    o = dev.DeviceArray.empty((8, 6, 1500, 64), dt)
    qd, kd, vd, od = q.desc(), k.desc(), v.desc(), o.desc()
    for _ in range(2): capi.check(L.axb_sdpa(by(qd), by(kd), by(vd), None, ctypes.c_float(-1.0), 0, by(od), None))
a = dev.randn((12000, 384), 21); w = dev.randn((1536, 384), 22, 0.0, 0.05); b = dev.randn((1536,), 23)
for _ in range(3): dev.matmul(a, w, transpose_b=True, bias=b, activation="gelu")   # gemm_tc_kernel (fp32 split)
w2 = dev.randn((384, 1536), 24, 0.0, 0.05); h = dev.randn((12000, 1536), 25)
for _ in range(3): dev.matmul(h, w2, transpose_b=True, bias=dev.randn((384,), 26))
dev.synchronize(); print("ok")
