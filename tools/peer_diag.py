"""Diagnostic for the single-process multi-rank mailbox test: per-call wall time and values."""
import ctypes, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from axiom_b200 import capi, device as dev, shard
L = capi.lib(); capi.check(L.axb_set_device(0))
rt = ctypes.CDLL("libcudart.so.12")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
streams = []
for _ in range(world):
    s = ctypes.c_void_p(); assert rt.cudaStreamCreateWithFlags(ctypes.byref(s), 1) == 0; streams.append(s)
for op in ("sum", "max", "min"):
    boxes = [shard.PeerGroup.create_mailbox()[0] for _ in range(world)]
    groups = [shard.PeerGroup(r, world, boxes) for r in range(world)]
    rng = np.random.default_rng(5)
    sizes = [(257, 1031), (64, 4096), (3, 70001), (1, 17)][:world]
    xs = [dev.DeviceArray.from_numpy(rng.standard_normal(sz).astype(np.float32) + np.float32(r)) for r, sz in enumerate(sizes)]
    outs = [dev.DeviceArray.empty((1,), capi.F32) for _ in range(world)]
    held = []
    for s in streams:
        for size in range(512, 8192 + 1, 512):
            ptr = ctypes.c_void_p(); capi.check(L.axb_alloc(size, s, ctypes.byref(ptr))); held.append((ptr, s))
    for ptr, s in held: capi.check(L.axb_free(ptr, s))
    ax = (ctypes.c_int32 * 2)(0, 1)
    for r in range(world):
        capi.check(L.axb_reduce(capi.OP[op], ctypes.byref(xs[r].desc()), ax, 2, 0, ctypes.byref(outs[r].desc()), streams[r]))
    for s in streams: rt.cudaStreamSynchronize(s)
    dev.synchronize()
    total = sum(int(np.prod(sz)) for sz in sizes)
    for call in range(4):
        t0 = time.perf_counter(); enq = []
        for r in range(world):
            groups[r].reduce(op, xs[r], total, outs[r], streams[r]); enq.append(round((time.perf_counter() - t0) * 1e3, 2))
        for s in streams: rt.cudaStreamSynchronize(s)
        t1 = time.perf_counter()
        print(op, "call", call, "enqueue ms", enq, "total ms", round((t1 - t0) * 1e3, 2), [float(o.numpy()[0]) for o in outs], flush=True)
