"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck); not a test of values."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
rng = np.random.default_rng(0)
f = lambda *s: dev.DeviceArray.from_numpy(rng.standard_normal(s).astype(np.float32))
by = ctypes.byref
# flash attention: fp32 three-piece and bf16, ragged lengths, causal, explicit mask
for dt in (capi.F32, capi.BF16):
    q, k, v = (f(2, 3, 301, 64) for _ in range(3))
    if dt == capi.BF16: q, k, v = (dev.cast(t, dt) for t in (q, k, v))
    o = dev.DeviceArray.empty((2, 3, 301, 64), dt)
    for causal in (0, 1):
        capi.check(L.axb_sdpa(by(q.desc()), by(k.desc()), by(v.desc()), None, ctypes.c_float(-1.0), causal, by(o.desc()), None))
# implicit-GEMM convolution: fp32 grouped stages, bf16 plain stages, strides / dilation / ragged channels, 1-D with gelu
dev.conv(f(2, 48, 40, 70), f(130, 48, 5, 3), f(130), stride=(1, 2), padding=(2, 0), dilation=(1, 1)); assert L.axb_matmul_last_path() == 4
dev.conv(f(2, 64, 57, 61), f(96, 64, 3, 3), f(96), stride=(2, 2), padding=(1, 1), dilation=(1, 1)); assert L.axb_matmul_last_path() == 4
xb, wb = dev.cast(f(2, 80, 33, 64), capi.BF16), dev.cast(f(64, 80, 3, 3), capi.BF16)
dev.conv(xb, wb, None, stride=(1, 1), padding=(2, 2), dilation=(2, 2)); assert L.axb_matmul_last_path() == 4
dev.conv(f(2, 80, 3000), f(384, 80, 3), f(384), stride=1, padding=1, dilation=1, activation="gelu"); assert L.axb_matmul_last_path() == 4
dev.conv(f(2, 384, 1001), f(384, 384, 3), None, stride=2, padding=1, dilation=1); assert L.axb_matmul_last_path() == 4
# half-precision table path (table build + lookup kernel, ragged tail)
xh = dev.cast(f((1 << 21) + 13), capi.BF16); dev.unary("gelu", xh); dev.unary("tanh", xh)
# pooling / sort (one-CTA and scratch paths, 32- and 64-bit keys) / topk / masked_select
x4 = f(2, 3, 13, 17); dev.pool(x4, "max", (3, 2), (2, 1), (1, 1)); dev.pool(x4, "avg", (3, 3), (2, 2), (1, 1), False)
dev.pool(f(1, 2, 6, 7, 9), "avg", (2, 3, 2), (1, 2, 2), (1, 1, 0))
dev.sort(f(5, 301, 3), 1, True); dev.sort(f(3, 4097, 2), 1, False); dev.sort(f(2, 70000), 1, True, k=17)
dev.sort(dev.DeviceArray.from_numpy(rng.integers(-9, 9, (3, 5000)).astype(np.int64)), 1, False)
dev.sort(dev.DeviceArray.from_numpy(rng.standard_normal((7, 300)).astype(np.float64)), 0, True)
m = dev.DeviceArray.from_numpy(rng.standard_normal((5, 33, 129)) > 0.2); dev.masked_select(f(5, 33, 129), m)
dev.masked_select(f(5, 33, 129).view((5, 129, 33), (33 * 129 * 4, 4, 129 * 4)), dev.DeviceArray.from_numpy(rng.standard_normal((129, 1)) > 0))
# chain + last-axis reduction: CTA-per-row and warp-per-row kernels, broadcast operands, ragged tails
F_, V_ = capi.F32, capi.SRC_VALUE
pr = [capi.make_instr(capi.OP["add"], F_, 0, 1), capi.make_instr(capi.OP["mul"], F_, V_(0), 2), capi.make_instr(capi.OP["relu"], F_, V_(1))]
for shp in ((37, 5000), (1000, 300), (4, 33, 1025), (5, 2)):
    for op in ("sum", "mean", "max"):
        dev.elementwise_reduce_axis(pr, [f(*shp[:-1], 1), f(shp[-1]), f(*shp)], op, -1)
# chain + softmax / log_softmax / layer_norm / rms_norm in one kernel (register-resident rows), with row epilogue
pr2 = [capi.make_instr(capi.OP["mul"], F_, 0, capi.SRC_IMM, imm=0.5), capi.make_instr(capi.OP["add"], F_, V_(0), 1)]
for shp in ((37, 5000), (1000, 300), (5, 2), (9, 8192)):
    ins = [f(*shp), f(shp[-1])]
    dev.elementwise_softmax(pr2, ins); dev.elementwise_softmax(pr2, ins, log=True); dev.elementwise_softmax(pr2, ins, epilogue="sum")
    dev.elementwise_norm(pr2, ins, "layer_norm", f(shp[-1]), f(shp[-1]), 1e-5); dev.elementwise_norm(pr2, ins, "rms_norm", f(shp[-1]), None, 1e-5)
# direct NCHW convolution (1x1, half precision) and multi-row tiles with a ragged last group
x1, w1 = dev.cast(f(3, 200, 27, 40), capi.BF16), dev.cast(f(256, 200, 1, 1), capi.BF16)
dev.conv(x1, w1, None, stride=(1, 1), padding=(0, 0), dilation=(1, 1)); assert L.axb_matmul_last_path() == 4
dev.conv(f(8, 64, 3, 64), f(64, 64, 3, 3), None, stride=(1, 1), padding=(1, 1), dilation=(1, 1)); assert L.axb_matmul_last_path() == 4
# staged pageable copies (six host threads, pinned lanes)
big = np.arange((40 << 20) // 4 + 3, dtype=np.int32); assert np.array_equal(dev.DeviceArray.from_numpy(big).numpy(), big)
dev.synchronize()
print("sanitize targets ok")
