"""Quick device-timed probe of individual workloads (dev tool; bench.py is the contract)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev

L = capi.lib()
capi.check(L.axb_set_device(0))

def timeit(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    capi.check(L.axb_event_create(ctypes.byref(e0))); capi.check(L.axb_event_create(ctypes.byref(e1)))
    ts = []
    for _ in range(iters):
        capi.check(L.axb_event_record(e0, None)); fn(); capi.check(L.axb_event_record(e1, None))
        capi.check(L.axb_event_sync(e1)); ms = ctypes.c_float(); capi.check(L.axb_event_elapsed_ms(e0, e1, ctypes.byref(ms))); ts.append(ms.value)
    return float(np.median(ts)), float(np.min(ts))

def c2(rows=4096, cols=16384):
    a = dev.randn((rows, 1), 1); b = dev.randn((1, cols), 2); c = dev.randn((rows, cols), 3)
    one = dev.DeviceArray.from_numpy(np.array([1.0], np.float32)); zi = dev.DeviceArray.from_numpy(np.array([0], np.int32)); zf = dev.DeviceArray.from_numpy(np.array([0.0], np.float32))
    F, B, V = capi.F32, capi.BOOL, capi.SRC_VALUE
    prog = [capi.make_instr(capi.OP["relu"], F, 0), capi.make_instr(capi.OP["add"], F, V(0), 1), capi.make_instr(capi.OP["mul"], F, V(1), 2),
            capi.make_instr(capi.OP["gt"], B, 3, 4), capi.make_instr(capi.OP["where"], F, V(3), V(2), 5)]
    out = dev.DeviceArray.empty((rows, cols), F)
    ins = [a, one, b, c, zi, zf]
    fn = lambda: dev.elementwise(prog, ins, out=out)
    med, mn = timeit(fn)
    nbytes = rows*cols*8 + rows*4 + cols*4
    return dict(workload="c2_chain", ms_med=med, ms_min=mn, GBps_med=nbytes/med/1e6, GBps_min=nbytes/mn/1e6)

def simple(op, n=1<<28):
    x = dev.randn((n,), 5); out = dev.DeviceArray.empty((n,), capi.F32)
    o = capi.OP[op]
    fn = lambda: capi.check(L.axb_unary(o, ctypes.byref(x.desc()), ctypes.byref(out.desc()), 0.0, None))
    med, mn = timeit(fn)
    return dict(workload="unary_"+op, ms_med=med, GBps_med=n*8/med/1e6, GBps_min=n*8/mn/1e6)

def simple_bf16(op, n=1<<28):
    x = dev.cast(dev.randn((n,), 5), capi.BF16); out = dev.DeviceArray.empty((n,), capi.BF16)
    o = capi.OP[op]
    xd, od = x.desc(), out.desc()
    fn = lambda: capi.check(L.axb_unary(o, ctypes.byref(xd), ctypes.byref(od), 0.0, None))
    med, mn = timeit(fn)
    return dict(workload="unary_bf16_"+op, ms_med=med, GBps_med=n*4/med/1e6, GBps_min=n*4/mn/1e6)

def add(n=1<<27):
    x = dev.randn((n,), 5); y = dev.randn((n,), 6); out = dev.DeviceArray.empty((n,), capi.F32)
    fn = lambda: capi.check(L.axb_binary(0, ctypes.byref(x.desc()), ctypes.byref(y.desc()), ctypes.byref(out.desc()), None))
    med, mn = timeit(fn)
    return dict(workload="add", ms_med=med, GBps_med=n*12/med/1e6, GBps_min=n*12/mn/1e6)

def d2d(n=1<<28):
    x = dev.randn((n,), 5); out = dev.DeviceArray.empty((n,), capi.F32)
    fn = lambda: capi.check(L.axb_memcpy_d2d(ctypes.c_void_p(out.storage.ptr), ctypes.c_void_p(x.storage.ptr), n*4, None))
    med, mn = timeit(fn)
    return dict(workload="cudaMemcpy_d2d", ms_med=med, GBps_med=n*8/med/1e6, GBps_min=n*8/mn/1e6)

def bmm(nb=64, m=4096, iters=5):
    A = dev.cast(dev.randn((nb, m, m), 11), capi.BF16); B = dev.cast(dev.randn((nb, m, m), 12, 0.0, 1.0/64), capi.BF16)
    C = dev.DeviceArray.empty((nb, m, m), capi.BF16)
    ad, bd, cd = A.desc(), B.desc(), C.desc()
    fn = lambda: capi.check(L.axb_matmul(ctypes.byref(ad), ctypes.byref(bd), 0, 0, None, ctypes.byref(cd), None))
    med, mn = timeit(fn, iters=iters, warm=2)
    fl = 2.0*nb*m**3
    return dict(workload=f"bmm_bf16_{nb}x{m}", ms_med=med, TFLOPs_med=fl/med/1e9, TFLOPs_best=fl/mn/1e9)

def bmm_nt(nb=64, m=4096, iters=5):
    A = dev.cast(dev.randn((nb, m, m), 11), capi.BF16); B = dev.cast(dev.randn((nb, m, m), 12, 0.0, 1.0/64), capi.BF16)
    C = dev.DeviceArray.empty((nb, m, m), capi.BF16)
    ad, bd, cd = A.desc(), B.desc(), C.desc()
    fn = lambda: capi.check(L.axb_matmul(ctypes.byref(ad), ctypes.byref(bd), 0, 1, None, ctypes.byref(cd), None))
    med, mn = timeit(fn, iters=iters, warm=2)
    fl = 2.0*nb*m**3
    return dict(workload=f"bmm_nt_bf16_{nb}x{m}", ms_med=med, TFLOPs_med=fl/med/1e9, TFLOPs_best=fl/mn/1e9)

def c3(kind, dt, rows=8192, n=32768):
    x = dev.randn((rows, n), 7); isz = 4
    if dt == "bf16":
        x = dev.cast(x, capi.BF16); isz = 2
    w, b = dev.randn((n,), 8, 1.0, 0.1), dev.randn((n,), 9, 0.0, 0.1)
    by = ctypes.byref
    xd, wd, bd = x.desc(), w.desc(), b.desc()
    ax = (ctypes.c_int32 * 1)(1)
    if kind in ("sum", "max"):
        o = dev.DeviceArray.empty((rows,), x.dtype); od = o.desc(); nb = rows * n * isz + rows * isz
        fn = lambda: capi.check(L.axb_reduce(capi.OP[kind], by(xd), ax, 1, 0, by(od), None))
    elif kind == "argmax":
        o = dev.DeviceArray.empty((rows,), capi.I64); od = o.desc(); nb = rows * n * isz + rows * 8
        fn = lambda: capi.check(L.axb_argreduce(capi.OP["argmax"], by(xd), 1, 0, by(od), None))
    else:
        o = dev.DeviceArray.empty((rows, n), x.dtype); od = o.desc(); nb = 2 * rows * n * isz
        if kind == "softmax":
            fn = lambda: capi.check(L.axb_softmax(by(xd), -1, 0, by(od), None))
        elif kind == "log_softmax":
            fn = lambda: capi.check(L.axb_softmax(by(xd), -1, 1, by(od), None))
        elif kind == "layer_norm":
            fn = lambda: capi.check(L.axb_layer_norm(by(xd), by(wd), by(bd), -1, 1e-5, by(od), None))
        elif kind == "layer_norm_noaffine":
            fn = lambda: capi.check(L.axb_layer_norm(by(xd), None, None, -1, 1e-5, by(od), None))
        else:
            fn = lambda: capi.check(L.axb_rms_norm(by(xd), by(wd), -1, 1e-5, by(od), None))
    med, mn = timeit(fn, iters=10, warm=3)
    return dict(workload=f"c3_{kind}_{dt}_{rows}x{n}", ms_med=med, GBps_med=nb/med/1e6, GBps_best=nb/mn/1e6)

def mm(dt, M, N, K, bias=0, act=None, iters=10):
    a = dev.randn((M, K), 21); w = dev.randn((N, K), 22, 0.0, 0.05)
    bvec = dev.randn((N,), 23) if bias else None
    if dt == "bf16":
        a = dev.cast(a, capi.BF16); w = dev.cast(w, capi.BF16); bvec = dev.cast(bvec, capi.BF16) if bias else None
    fn = lambda: dev.matmul(a, w, transpose_b=True, bias=bvec, activation=act)
    med, mn = timeit(fn, iters=iters, warm=3)
    fl = 2.0 * M * N * K
    return dict(workload=f"linear_{dt}_{M}x{N}x{K}_bias{bias}_{act}", ms_med=med, TFLOPs_med=fl/med/1e9, TFLOPs_best=fl/mn/1e9)

def conv(dt, xs, ws, stride, padding, act=None, iters=10):
    """conv:<dt>:<N,C,H,W>:<Cout,Cin,KH,KW>:<stride>:<padding>[:act] — 1-D when the shapes are 3-D"""
    x = dev.randn(tuple(xs), 31); w = dev.randn(tuple(ws), 32, 0.0, 0.05); b = dev.randn((ws[0],), 33)
    if dt != "f32":
        code = capi.BF16 if dt == "bf16" else capi.F16
        x, w, b = dev.cast(x, code), dev.cast(w, code), dev.cast(b, code)
    nsp = len(xs) - 2
    fn = lambda: dev.conv(x, w, b, stride=stride, padding=padding, dilation=1, activation=act)
    out = fn()
    path = L.axb_matmul_last_path()
    med, mn = timeit(fn, iters=iters, warm=3)
    fl = 2.0 * np.prod(out.shape) * ws[1] * np.prod(ws[2:])
    return dict(workload=f"conv{nsp}d_{dt}_{'x'.join(map(str, xs))}_{'x'.join(map(str, ws))}_s{stride}_p{padding}_{act}", path=int(path),
                ms_med=med, TFLOPs_med=fl/med/1e9, TFLOPs_best=fl/mn/1e9)

if __name__ == "__main__":
    which = sys.argv[1:] or ["d2d", "relu", "exp", "gelu", "add", "c2"]
    for w in which:
        if w.startswith("c3:"):
            _, kind, dt = w.split(":")[:3]
            extra = [int(v) for v in w.split(":")[3:]]
            print(json.dumps(c3(kind, dt, *extra)), flush=True)
            continue
        if w.startswith("bf16:"):
            print(json.dumps(simple_bf16(w.split(":")[1])), flush=True)
            continue
        if w.startswith("conv:"):
            f = w.split(":")
            print(json.dumps(conv(f[1], [int(v) for v in f[2].split(",")], [int(v) for v in f[3].split(",")], int(f[4]), int(f[5]),
                                  f[6] if len(f) > 6 else None)), flush=True)
            continue
        if w.startswith("mm:"):
            f = w.split(":")
            print(json.dumps(mm(f[1], int(f[2]), int(f[3]), int(f[4]), int(f[5]) if len(f) > 5 else 0, f[6] if len(f) > 6 else None)), flush=True)
            continue
        r = bmm() if w == "bmm" else bmm_nt() if w == "bmm_nt" else bmm(8, 4096) if w == "bmm8" else d2d() if w == "d2d" else c2() if w == "c2" else add() if w == "add" else simple(w)
        print(json.dumps(r), flush=True)
