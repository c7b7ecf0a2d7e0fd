"""Worker of tests/test_multi_gpu.py: one process per GPU under torchrun (NCCL). Exercises the REAL cross-GPU paths of
SURVEY.md §8(e): axb_reduce_sharded (peer-memory mailboxes over NVLink, fused into the reduction's finish kernel) and
shard.combine_argreduce, on ragged shards, fp32 and bf16, back-to-back without host synchronisation, plus the
"validation failure on one rank must not desynchronise the call counter" rule. Rank 0 prints one JSON line."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from axiom_b200 import capi, device as dev, shard
    from oracle import axiom_oracle as ao

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    L = capi.lib()
    capi.check(L.axb_set_device(local))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    devname = f"cuda:{local}"
    res = {"world": world, "cases": 0, "failures": []}

    def fail(msg):
        res["failures"].append(f"rank {rank}: {msg}")

    # every rank can rebuild every shard (seeded by shard index): ragged row blocks of a {rows, cols} tensor
    rows, cols = 2 * 1000 * world + 3, 4096
    blocks = [shard.plan_rows(rows, world, r) for r in range(world)]

    def shard_np(r):
        lo, hi = blocks[r]
        x = np.random.default_rng(1000 + r).standard_normal((hi - lo, cols)).astype(np.float32)
        return x

    full = np.concatenate([shard_np(r) for r in range(world)], axis=0)
    mine = shard_np(rank)
    group = shard.PeerGroup.from_process_group(devname)
    try:
        for dt in ("f32", "bf16"):
            if dt == "f32":
                x = dev.DeviceArray.from_numpy(mine)
                ref_full = full.astype(np.float64)
            else:
                bits = ao.f32_to_bf16(mine)
                x = dev.DeviceArray.from_numpy(bits, bf16=True)
                ref_full = ao.bf16_to_f32(ao.f32_to_bf16(full)).astype(np.float64)
            for op, f in (("sum", np.sum), ("mean", np.mean), ("max", np.max), ("min", np.min)):
                out_dt = capi.F32 if (dt == "f32" or op == "mean") else capi.BF16
                out = dev.DeviceArray.empty((1,), out_dt)
                group.reduce(op, x, rows * cols, out)
                got = out.numpy()
                got = float(got[0]) if out_dt == capi.F32 else float(ao.bf16_to_f32(got)[0])
                exp = float(f(ref_full))
                tol = (1e-5 * float(np.abs(ref_full).sum()) if op == "sum" else 1e-5 * float(np.abs(ref_full).mean()) + 1e-7) \
                    if op in ("sum", "mean") else 0.0
                if dt == "bf16" and op in ("sum",):
                    tol = max(tol, abs(exp) * 2.0 ** -8)
                if abs(got - exp) > tol:
                    fail(f"{op} {dt}: got {got}, expected {exp} (tol {tol})")
                # identical on all ranks (rank-ordered combine)
                t = torch.tensor([got], dtype=torch.float64, device=devname)
                allv = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allv, t)
                if len({float(v.item()) for v in allv}) != 1:
                    fail(f"{op} {dt}: ranks disagree: {[float(v.item()) for v in allv]}")
                res["cases"] += 1
        # 200 exchanges enqueued back to back, no host synchronisation in between (two-bank mailbox reuse)
        x = dev.DeviceArray.from_numpy(mine)
        outs = [dev.DeviceArray.empty((1,), capi.F32) for _ in range(200)]
        for o in outs:
            group.reduce("sum", x, rows * cols, o)
        dev.synchronize()
        vals = {float(o.numpy()[0]) for o in outs}
        if len(vals) != 1:
            fail(f"back-to-back exchanges returned {len(vals)} distinct values")
        res["cases"] += 1
        # a validation failure on ONE rank must not move its call counter: rank 0 issues a rejected call
        if rank == 0:
            bad = dev.DeviceArray.from_numpy(np.arange(16, dtype=np.int32))
            try:
                group.reduce("sum", bad, 16 * world, dev.DeviceArray.empty((1,), capi.I32))
                fail("int32 sharded reduction was not rejected")
            except capi.AxbError:
                pass
        out = dev.DeviceArray.empty((1,), capi.F32)
        group.reduce("max", x, rows * cols, out)
        if float(out.numpy()[0]) != float(full.max()):
            fail("exchange after a rejected call is out of step")
        group.status()
        res["cases"] += 1
        # full argmax of the row-sharded tensor: local (value, index) on the GPU, combine over the process group
        tie = full.copy()
        tie[blocks[-1][0] + 5, 17] = 50.0
        tie[3, 9] = 50.0  # the same maximum in the first and the last shard: the lowest flat index wins
        tie[1, 1] = np.nan
        lo, hi = blocks[rank]
        xs = dev.DeviceArray.from_numpy(np.ascontiguousarray(tie[lo:hi]))
        li = int(dev.argreduce("argmax", xs, -1).numpy().reshape(-1)[0])
        lv = float(tie[lo:hi].reshape(-1)[li])
        gi = shard.combine_argreduce("argmax", lv, li, lo * cols, device=devname)
        exp_i = int(np.argmax(np.where(np.isnan(tie), -np.inf, tie).reshape(-1)))
        if gi != exp_i:
            fail(f"sharded argmax: got {gi}, expected {exp_i}")
        res["cases"] += 1
    finally:
        dev.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        group.close()
    allf = [None] * world
    dist.all_gather_object(allf, res["failures"])
    if rank == 0:
        res["failures"] = [m for fl in allf for m in fl]
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
