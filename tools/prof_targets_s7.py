"""ncu targets added in session 7: fused chain + full reduction, bf16 gelu (fp32 evaluation + double fallback)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from axiom_b200 import capi, device as dev
L = capi.lib(); capi.check(L.axb_set_device(0))
F, V = capi.F32, capi.SRC_VALUE
n = 1 << 26
a, b, c = dev.randn((n,), 1), dev.randn((n,), 2), dev.randn((n,), 3)
prog = [capi.make_instr(capi.OP["add"], F, 0, 1), capi.make_instr(capi.OP["mul"], F, V(0), 2)]
for _ in range(2): dev.elementwise_reduce(prog, [a, b, c], "sum")
x = dev.cast(dev.randn((1 << 28,), 5), capi.BF16)
for op in ("relu", "gelu", "exp"):
    for _ in range(2): dev.unary(op, x)
dev.synchronize(); print("ok")
