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
# CTA-pair GEMM (153 tiles of 256x256, ragged M / N, one k-block and three): cluster barriers, remote arrives, multicast commits
ap = dev.cast(f(2250, 64), capi.BF16); bp = dev.cast(f(64, 4330 // 8 * 8), capi.BF16); dev.matmul(ap, bp); assert capi.lib().axb_matmul_last_path() == 2
ap = dev.cast(f(2304, 192), capi.BF16); bp = dev.cast(f(4352, 192), capi.BF16); dev.matmul(ap, bp, transpose_b=True, bias=dev.cast(f(4352), capi.BF16), activation="relu")
# half-type transcendentals (fp32 evaluation + double fallback), bf16 pair rounding / PRMT store, ragged tail
xh = dev.cast(f(1000, 37), capi.BF16); dev.unary("gelu", xh); dev.unary("exp", xh); dev.unary("relu", xh)
# spatial operators
xs = f(2, 4, 9, 10); dev.conv2d(xs, f(8, 2, 3, 2), f(8), stride=(2, 1), padding=(1, 2), dilation=(1, 2), groups=2)
dev.conv_transpose(f(2, 4, 5, 6), f(4, 3, 3, 2), f(6), stride=(2, 3), padding=(1, 0), output_padding=(1, 2), groups=2)
dev.conv_transpose(f(3, 4, 11), f(4, 2, 5), None, stride=3, padding=2, output_padding=1, dilation=2, groups=2)
for m in ("nearest", "bilinear", "bicubic"): dev.interpolate(xs, (7, 13), m, True)
# fused chain + full reduction (EW_RED: per-thread fold, CTA reduce, one partial per CTA), ragged sizes, broadcast column
F_, V_ = capi.F32, capi.SRC_VALUE
pr = [capi.make_instr(capi.OP["add"], F_, 0, 1), capi.make_instr(capi.OP["mul"], F_, V_(0), 2), capi.make_instr(capi.OP["relu"], F_, V_(1))]
for shp in ((257, 1031), (1, 5), (3, 70001)):
    for op in ("sum", "mean", "max"): dev.elementwise_reduce(pr, [f(shp[0], 1), f(*shp), f(*shp)], op)
dev.synchronize()
print("sanitize targets ok")
