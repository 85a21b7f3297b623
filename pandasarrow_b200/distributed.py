"""Multi-GPU group-by: one process per GPU, rows sharded by contiguous range.

    local aggregate (stages 1-3 on this rank's shard)
      -> partial records bucketed by owner = hash(key) % world   (CUDA, pa_groupby_partials_*)
      -> all-to-all of counts, then of records                   (torch.distributed: NCCL over NVLink)
      -> merge by key in source-rank order + order by global first row   (CUDA, pa_merge_create)

Each rank ends up owning the groups with hash(key) % world == rank; `gather_result` concatenates
the owners' results on every rank and restores global first-appearance order.

torch.distributed is plumbing only (rendezvous + the collective); the exchange helpers work on any
backend, which is what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import PA_PARTIAL_WORDS

REC_KEY, REC_FLAGS, REC_SUM, REC_DSUM, REC_COUNT, REC_FIRST_ROW = 0, 1, 2, 3, 4, 5


def shard_rows(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [begin, end) of `rank` (rank r owns rows [r*N/P, (r+1)*N/P))."""
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def _mix64(k: np.ndarray) -> np.ndarray:
    """Host twin of hash_key64 (csrc/common.cuh)."""
    k = k.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xFF51AFD7ED558CCD)
        k ^= k >> np.uint64(33)
        k *= np.uint64(0xC4CEB9FE1A85EC53)
        k ^= k >> np.uint64(33)
    return k


def owner_of(keys: np.ndarray, world: int, is_null: Optional[np.ndarray] = None) -> np.ndarray:
    """Owner rank of every key (null keys live on rank 0) — must match owner_of() in csrc/merge.cuh."""
    o = (_mix64(np.asarray(keys).view(np.uint64) if np.asarray(keys).dtype.itemsize == 8 else np.asarray(keys)) %
         np.uint64(world)).astype(np.int64)
    if is_null is not None:
        o[np.asarray(is_null, dtype=bool)] = 0
    return o


def exchange_records(send, send_counts: Sequence[int], group=None):
    """All-to-all of fixed-size records.  `send`: int64 tensor [sum(send_counts), PA_PARTIAL_WORDS],
    grouped by destination rank.  Returns (recv tensor grouped by source rank, recv_counts)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert len(send_counts) == world
    dev = send.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.tolist()]
    recv = torch.empty((sum(recv_counts), PA_PARTIAL_WORDS), dtype=torch.int64, device=dev)
    w = PA_PARTIAL_WORDS
    dist.all_to_all_single(recv.view(-1), send.reshape(-1), output_split_sizes=[c * w for c in recv_counts],
                           input_split_sizes=[int(c) * w for c in send_counts], group=group)
    return recv, recv_counts


class Comm:
    """pa_comm (include/pa_b200.h): an NCCL communicator owned by libpa_b200.so.  The unique id travels over the
    already initialised torch.distributed group (any backend); after that the whole sharded step — local pass, owner
    bucketing, count / record exchange, merge — is ONE C call (pa_groupby_sharded_aggregate), and this module is
    just its caller."""

    def __init__(self, group=None, device: Optional[int] = None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        from .groupby import _check
        L = _lib.load()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.cuda.current_device() if device is None else device
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            _check(L.pa_comm_unique_id(buf, 128))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0, group=group)
        raw = bytes(t.cpu().tolist())
        self._h = C.c_void_p()
        self.world, self.rank, self.device = world, rank, dev
        _check(L.pa_comm_create(raw, world, rank, dev, C.byref(self._h)))
        self._L = L

    def sharded_aggregate(self, gb, values, aggs: Sequence[str]):
        """One multi-GPU step; returns the owner-side MergedGroupBy."""
        import ctypes as C
        from ._lib import PA_AGG
        from .groupby import MergedGroupBy, _CArg, _check
        mask = 0
        for a in aggs:
            mask |= PA_AGG[a]
        arg = _CArg(values)
        h = C.c_void_p()
        try:
            _check(self._L.pa_groupby_sharded_aggregate(gb._h, self._h, C.byref(arg.dev), C.byref(arg.schema), mask, C.byref(h)))
        finally:
            arg.close()
        return MergedGroupBy._from_handle(h)

    def phases(self) -> Dict[str, float]:
        import ctypes as C
        from .groupby import _check
        p = (C.c_double * 5)()
        _check(self._L.pa_comm_last_phases(self._h, p))
        return {"local_ms": p[0], "export_ms": p[1], "exchange_ms": p[2], "merge_ms": p[3], "total_ms": p[4]}

    def exchange_info(self) -> Dict[str, int]:
        """Record format of the last step (pa_comm_last_exchange)."""
        import ctypes as C
        from .groupby import _check
        rb, un, slots = C.c_int32(), C.c_int32(), C.c_int64()
        _check(self._L.pa_comm_last_exchange(self._h, C.byref(rb), C.byref(un), C.byref(slots)))
        return {"record_bytes": rb.value, "unordered_export": bool(un.value), "merge_table_slots": slots.value}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.pa_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PADDED_BLOCK_RECORDS = 2048   # padded exchange: up to this many groups per rank (180 KB per peer block)


def exchange_padded(send_blocks, group=None):
    """Equal-split all-to-all of fixed-size blocks: send_blocks[p] goes to rank p, the result's block s came
    from rank s.  No counts are exchanged and nothing is read back: the whole step stays stream ordered."""
    import torch
    import torch.distributed as dist
    recv = torch.empty_like(send_blocks)
    dist.all_to_all_single(recv.view(-1), send_blocks.view(-1), group=group)
    return recv


def sharded_aggregate(gb, values, aggs: Sequence[str], value_format: str, key_format: str, group=None, stream=None,
                      padded: Optional[bool] = None, wait: bool = True, tail_stream=None):
    """Run the whole multi-GPU step for this rank.  `gb` is this rank's GroupBy over its row shard
    (created with row_base = first global row of the shard).  Returns the owner-side MergedGroupBy.

    With few groups (<= PADDED_BLOCK_RECORDS on this rank) the partials travel in fixed-size blocks: one
    stream-ordered export kernel, one equal-split all-to-all, one merge — no host round trip in between.  A rank
    with more groups sends an overflow marker to every peer, so all ranks fall back to the counted exchange
    together.  `stream` must be the current torch stream for the padded path (it is in bench.py).

    wait=False (padded path only): nothing is read back at all — local pass, export, exchange and merge are just
    queued, so consecutive steps pipeline; the first call on the returned handle that needs a result completes
    it (and raises PaError if a rank had overflowed: rerun that step with wait=True).

    tail_stream (a torch.cuda.Stream, padded path only): the exchange and the merge run on it, behind an event recorded
    after the export — the handle's own stream is free for the next step's local pass while this step's records travel
    and merge (the serial tail is launch / NCCL latency on a handful of CTAs)."""
    import torch
    import torch.distributed as dist
    from .groupby import MergedGroupBy, PaError
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if padded is None:
        padded = stream is not None and stream == torch.cuda.current_stream(dev).cuda_stream
    # padded path: the local pass is only queued (no status read-back); export, exchange and merge are queued
    # behind it, so the host never waits before the merge's own result read.  A local pass that did not fit
    # the optimistic path makes the export send overflow markers, and everything is redone synchronously below.
    gb.aggregate(values, aggs, fetch=False, wait=not padded)
    if padded:
        cap = PADDED_BLOCK_RECORDS
        send = torch.empty((world, cap + 1, PA_PARTIAL_WORDS), dtype=torch.int64, device=dev)
        gb.partials_export_padded(world, send.data_ptr(), cap)
        try:
            if tail_stream is None:
                recv = exchange_padded(send, group)
                merged = MergedGroupBy(recv.data_ptr(), [0] * world, aggs, value_format, key_format, device=dev.index,
                                       stream=stream, padded_block_records=cap)
            else:
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(dev))
                send.record_stream(tail_stream)
                with torch.cuda.stream(tail_stream):
                    tail_stream.wait_event(done)
                    recv = exchange_padded(send, group)
                    merged = MergedGroupBy(recv.data_ptr(), [0] * world, aggs, value_format, key_format, device=dev.index,
                                           stream=tail_stream.cuda_stream, padded_block_records=cap)
            merged._keep = (send, recv)
            if wait:
                merged.groupSize()          # completes the merge; raises if any rank sent the overflow marker
            return merged
        except PaError as e:
            if "padded block" not in str(e):
                raise
    counts = gb.partials_count(world)
    dev = torch.device("cuda", torch.cuda.current_device())
    send = torch.empty((max(sum(counts), 1), PA_PARTIAL_WORDS), dtype=torch.int64, device=dev)
    gb.partials_export(world, send.data_ptr(), send.shape[0])
    recv, recv_counts = exchange_records(send[:sum(counts)], counts, group)
    torch.cuda.synchronize()
    merged = MergedGroupBy(recv.data_ptr(), recv_counts, aggs, value_format, key_format, device=dev.index, stream=stream)
    merged._keep = (send, recv)
    return merged


def resample_anchor(first_ts_local: Optional[int], ticks_per_day: int = 86_400 * 10**9, group=None) -> int:
    """Common bucket anchor for a row-range sharded resample (last row): the reference anchors the
    bucket grid at the start of the day of the FIRST timestamp (TimeGrouperOrigin::StartDay, resample.cpp:85-178);
    every rank must use the global first timestamp, not its shard's.  Returns start_day(min over ranks) to pass as
    `origin="custom", origin_custom_ns=...`.  first_ts_local: this shard's first timestamp (None = empty shard)."""
    import torch
    import torch.distributed as dist
    big = (1 << 63) - 1
    t = torch.tensor([big if first_ts_local is None else int(first_ts_local)], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    first = int(t.item())
    if first == big:
        return 0
    return first - first % ticks_per_day          # (Python's % is a floor modulus: correct before the epoch too)


def gather_result(local: Dict[str, "np.ndarray"], first_rows: "np.ndarray", group=None) -> Dict[str, "np.ndarray"]:
    """All-gather the owners' result columns and restore global first-appearance order
    (stable sort by global first row).  Host-side convenience for tests / small results."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts: List = [None] * world
    dist.all_gather_object(parts, (local, first_rows), group=group)
    fr = np.concatenate([p[1] for p in parts])
    order = np.argsort(fr, kind="stable")
    return {k: np.concatenate([p[0][k] for p in parts])[order] for k in local}
