"""Batch-dimension sharding across the GPUs of one box.

The reference has no multi-device notion (enum Device {CPU, GPU} carries no index, include/axiom/storage.hpp:8-11), so
this is new functionality that lives BESIDE the unchanged API: one process per GPU, every rank owns a contiguous block of
the leading (batch / row) dimension, and the only exchange step is the combine of FULL-tensor reductions — per-row
reductions, softmax / norms, batched matmul and elementwise chains need no collective at all.

`torch.distributed` is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests); the partials themselves
come from libaxb kernels (or, in the CPU tests, from numpy). Nothing here computes on the data path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def plan_rows(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of `rank`'s block of a dimension of extent `total`: blocks differ by at most one row, earlier
    ranks take the extra ones (so shard r starts at r*q + min(r, rem))."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    q, rem = divmod(int(total), world)
    start = rank * q + min(rank, rem)
    return start, start + q + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def combine_full_reduction(op: str, partial: np.ndarray, local_count: int, device: Optional[str] = None) -> np.ndarray:
    """Combines per-shard partials of a full-tensor reduction with ONE all-reduce.

    sum/prod/max/min/any/all map onto the matching collective op; mean all-reduces (sum, count) so that ragged shards
    are weighted correctly. `partial` is the shard's {1} result in its accumulation dtype (fp32 for half inputs)."""
    import torch
    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(partial).copy())
    if device:
        t = t.to(device)
    if op == "mean":
        acc = torch.stack([t.double().reshape(-1)[0] * float(local_count), torch.tensor(float(local_count), dtype=torch.float64, device=t.device)])
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        out = (acc[0] / acc[1]).to(t.dtype).reshape(t.shape)
    else:
        red = {"sum": dist.ReduceOp.SUM, "prod": dist.ReduceOp.PRODUCT, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN,
               "any": dist.ReduceOp.MAX, "all": dist.ReduceOp.MIN}[op]
        dist.all_reduce(t, op=red)
        out = t
    return out.cpu().numpy()


def combine_argreduce(op: str, value: float, local_index: int, shard_offset: int, device: Optional[str] = None) -> int:
    """Full argmax/argmin of a row-sharded flattened tensor: all-gather (value, global index) pairs and apply the
    reference's rule on the host — strict compare from lowest()/max(), first (lowest index) extreme wins, NaN never wins
    (cpu_operations.cpp:2211-2228). `shard_offset` = number of elements that precede this shard in C order."""
    import torch
    dist = _dist()
    pair = torch.tensor([float(value), float(local_index + shard_offset)], dtype=torch.float64)
    if device:
        pair = pair.to(device)
    gathered = [torch.empty_like(pair) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, pair)
    best_val = -np.finfo(np.float64).max if op == "argmax" else np.finfo(np.float64).max
    best_idx = 0
    for g in sorted((g.cpu().numpy() for g in gathered), key=lambda a: a[1]):
        v, i = float(g[0]), int(g[1])
        if (op == "argmax" and v > best_val) or (op == "argmin" and v < best_val):
            best_val, best_idx = v, i
    return best_idx


class PeerGroup:
    """Mailboxes in peer memory for axb_reduce_sharded (include/axb.h): the per-shard partials of a full reduction are
    exchanged by the reduction's own finish kernel over NVLink instead of by a collective call. torch.distributed only
    carries the 64-byte IPC handles once, at construction (`local_boxes` builds a group from mailboxes of this process —
    used by the single-GPU test of the protocol)."""

    def __init__(self, rank: int, world: int, boxes, opened=(), own=None):
        import ctypes
        from . import capi
        self._capi, self.rank, self.world = capi, rank, world
        self._opened, self._own = list(opened), own
        arr = (ctypes.c_void_p * world)(*[ctypes.c_void_p(b) for b in boxes])
        g = ctypes.c_void_p()
        capi.check(capi.lib().axb_peer_group_create(rank, world, arr, ctypes.byref(g)))
        self.handle = g

    @staticmethod
    def create_mailbox():
        import ctypes
        from . import capi
        box, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        capi.check(capi.lib().axb_peer_mailbox_create(ctypes.byref(box), h))
        return box.value, h.raw

    @classmethod
    def from_process_group(cls, device: str) -> "PeerGroup":
        """One process per GPU: all-gather the IPC handles over the default process group and map every peer's mailbox.
        A failure on any rank (no peer access, IPC disabled) is agreed on collectively, so either every rank gets a group
        or every rank raises — nobody is left waiting in a collective."""
        import ctypes
        import torch
        from . import capi
        dist = _dist()
        rank, world = dist.get_rank(), dist.get_world_size()
        own, handle, err = None, bytes(64), ""
        try:
            own, handle = cls.create_mailbox()
        except Exception as ex:  # noqa: BLE001 — reported after the collective below
            err = repr(ex)
        mine = torch.tensor(list(handle) + [1 if own else 0], dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        rows = [g.cpu().tolist() for g in gathered]
        boxes, opened = [], []
        if all(r[64] == 1 for r in rows):
            try:
                for r in range(world):
                    if r == rank:
                        boxes.append(own)
                        continue
                    p = ctypes.c_void_p()
                    capi.check(capi.lib().axb_peer_mailbox_open(bytes(rows[r][:64]), ctypes.byref(p)))
                    boxes.append(p.value)
                    opened.append(p.value)
            except Exception as ex:  # noqa: BLE001
                err = repr(ex)
        else:
            err = err or "a peer could not create its mailbox"
        flag = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            L = capi.lib()
            for p in opened:
                L.axb_peer_mailbox_close(ctypes.c_void_p(p))
            if own:
                L.axb_peer_mailbox_destroy(ctypes.c_void_p(own))
            raise RuntimeError("peer mailboxes unavailable: " + (err or "failure on another rank"))
        return cls(rank, world, boxes, opened, own)

    def reduce(self, op: str, shard, global_count: int, out, stream=None) -> None:
        """Full `op` over all shards; `out` ({1}, device) receives the combined value on every rank."""
        import ctypes
        capi = self._capi
        capi.check(capi.lib().axb_reduce_sharded(capi.OP[op], ctypes.byref(shard.desc()), int(global_count), self.handle,
                                                 ctypes.byref(out.desc()), stream))

    def status(self) -> None:
        """Raises AxbError once any exchange of this group has timed out (a peer never delivered its partial)."""
        self._capi.check(self._capi.lib().axb_peer_group_status(self.handle))

    def close(self) -> None:
        import ctypes
        L = self._capi.lib()
        if self.handle:
            L.axb_peer_group_destroy(self.handle)
            self.handle = None
        for p in self._opened:
            L.axb_peer_mailbox_close(ctypes.c_void_p(p))
        self._opened = []
        if self._own:
            L.axb_peer_mailbox_destroy(ctypes.c_void_p(self._own))
            self._own = None
