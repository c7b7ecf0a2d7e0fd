"""Host-side logic of the multi-GPU path, exercised with world_size-2 gloo process groups on CPU:
row-block planning, the single all-reduce that combines full-tensor reductions, and the argmax pair combine."""
import os
import socket

import numpy as np
import pytest

from axiom_b200 import shard


def test_plan_rows_covers_everything_once():
    for total in (0, 1, 7, 64, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.plan_rows(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a1 - a0 >= b1 - b0 >= 0
            assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    with pytest.raises(ValueError):
        shard.plan_rows(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        x = rng.standard_normal((37, 50)).astype(np.float32)  # 37 rows: ragged split 19 + 18
        x[20, 3] = 9.0
        x[5, 7] = 9.0  # tie: the first (lowest flat index) must win
        x[30, 1] = np.nan
        lo, hi = shard.plan_rows(x.shape[0], world, rank)
        mine = x[lo:hi]
        clean = np.nan_to_num(mine, nan=0.0)
        out = {}
        for op, f in (("sum", np.sum), ("max", np.max), ("min", np.min)):
            out[op] = shard.combine_full_reduction(op, np.array([f(clean)], np.float32), clean.size)
        out["mean"] = shard.combine_full_reduction("mean", np.array([clean.mean()], np.float32), clean.size)
        out["any"] = shard.combine_full_reduction("any", np.array([float((clean > 8).any())], np.float32), clean.size)
        out["all"] = shard.combine_full_reduction("all", np.array([float((clean > -100).all())], np.float32), clean.size)
        flat = mine.reshape(-1)
        masked = np.where(np.isnan(flat), -np.inf, flat)
        li = int(np.argmax(masked))
        out["argmax"] = shard.combine_argreduce("argmax", float(flat[li]), li, lo * x.shape[1])
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_full_reduction_combine_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    x = rng.standard_normal((37, 50)).astype(np.float32)
    x[20, 3] = 9.0
    x[5, 7] = 9.0
    x[30, 1] = np.nan
    clean = np.nan_to_num(x, nan=0.0)
    for rank in (0, 1):
        r = results[rank]
        assert np.allclose(r["sum"], clean.sum(), rtol=1e-5)
        assert r["max"][0] == clean.max() and r["min"][0] == clean.min()
        assert np.allclose(r["mean"], clean.mean(), rtol=1e-5, atol=1e-7)
        assert r["any"][0] == 1.0 and r["all"][0] == 1.0
        assert r["argmax"] == 5 * 50 + 7  # first of the tied maxima, NaN ignored


def _peer_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        try:
            shard.PeerGroup.from_process_group("cpu")
            q.put((rank, "created"))
        except RuntimeError as ex:
            q.put((rank, "raised: " + str(ex)[:80]))
    finally:
        dist.destroy_process_group()


def test_peer_group_failure_is_agreed_on_by_all_ranks_gloo():
    """axb_reduce_sharded's mailboxes need a GPU. Without one, PeerGroup.from_process_group must fail on EVERY rank after
    the same collectives (handle all-gather, status all-reduce) — no rank may raise early and leave the other one waiting.
    With GPUs (not here) both ranks would report 'created'."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    kinds = {v.split(":")[0] for v in results.values()}
    assert len(kinds) == 1, results  # both created or both raised
    import torch
    if not torch.cuda.is_available():
        assert kinds == {"raised"}, results
