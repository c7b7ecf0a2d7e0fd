"""ncu targets of round 2, second batch: implicit-GEMM convolution (fp32 grouped / bf16 plain / direct NCHW), the fused row kernels
(chain + last-axis sum, chain + softmax, chain + layer_norm), pooling, sort, masked_select."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
F, V = capi.F32, capi.SRC_VALUE
cast = lambda t, c: dev.cast(t, c)
# conv
xb, wb = cast(dev.randn((32, 256, 56, 56), 1), capi.BF16), cast(dev.randn((256, 256, 3, 3), 2, 0.0, 0.02), capi.BF16)
for _ in range(2): dev.conv(xb, wb, None, stride=(1, 1), padding=(1, 1), dilation=(1, 1))
w1 = cast(dev.randn((512, 256, 1, 1), 3, 0.0, 0.05), capi.BF16)
for _ in range(2): dev.conv(xb, w1, None, stride=(1, 1), padding=(0, 0), dilation=(1, 1))
xf, wf = dev.randn((32, 64, 56, 56), 4), dev.randn((64, 64, 3, 3), 5, 0.0, 0.05)
for _ in range(2): dev.conv(xf, wf, None, stride=(1, 1), padding=(1, 1), dilation=(1, 1))
del xb, wb, w1, xf, wf
# fused row kernels
a, b = dev.randn((8192, 32768), 6), dev.randn((1, 32768), 7)
prog = [capi.make_instr(capi.OP["mul"], F, 0, 1)]
for _ in range(2): dev.elementwise_reduce_axis(prog, [a, b], "sum", 1)
del a, b
x, y = dev.randn((8192, 8192), 8), dev.randn((8192, 1), 9)
prog2 = [capi.make_instr(capi.OP["add"], F, 0, 1)]
for _ in range(2): dev.elementwise_softmax(prog2, [x, y])
for _ in range(2): dev.elementwise_norm(prog2, [x, y], "layer_norm", None, None, 1e-5)
# pooling / sort / masked_select
xp = dev.randn((32, 64, 112, 112), 10)
for _ in range(2): dev.pool(xp, "max", (3, 3), (2, 2), (1, 1))
xs = dev.randn((4096, 4096), 11)
for _ in range(2): dev.sort(xs, 1, False)
xl = dev.randn((8, 1 << 20), 12)
dev.sort(xl, 1, True, k=64)
m = dev.DeviceArray.from_numpy(np.random.default_rng(0).random((4096, 4096)) > 0.5)
dev.masked_select(xs, m)
dev.synchronize(); print("ok")
