"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exchange partial
records exactly as the NCCL ranks do (pandasarrow_b200/distributed.py).  No GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

W = 11


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_records(rank, world, n):
    """Records this rank would send: key k is owned by owner_of(k); payload encodes (rank, key)."""
    from pandasarrow_b200 import distributed as D
    rng = np.random.default_rng(100 + rank)
    keys = rng.integers(0, 500, n).astype(np.int64)
    keys = np.unique(keys)                                  # one record per key and source, like real partials
    own = D.owner_of(keys, world)
    order = np.argsort(own, kind="stable")
    keys, own = keys[order], own[order]
    rec = np.zeros((len(keys), W), dtype=np.int64)
    rec[:, 0] = keys
    rec[:, 2] = keys * 10 + rank                            # "sum"
    rec[:, 4] = 1 + rank                                    # "count"
    rec[:, 5] = keys * world + rank                         # "first_row" (distinct)
    counts = [int((own == o).sum()) for o in range(world)]
    return rec, counts


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pandasarrow_b200 import distributed as D
    rec, counts = _fake_records(rank, world, 300)
    recv, rc = D.exchange_records(torch.from_numpy(rec), counts)
    recv = recv.numpy()
    ok = True
    # everything received is owned by this rank, grouped by source rank, nothing lost
    ok &= bool((D.owner_of(recv[:, 0], world) == rank).all())
    off = 0
    for s in range(world):
        seg = recv[off:off + rc[s]]
        ok &= bool(((seg[:, 2] - seg[:, 0] * 10) == s).all())
        exp_rec, exp_counts = _fake_records(s, world, 300)
        e_off = sum(exp_counts[:rank])
        ok &= np.array_equal(seg, exp_rec[e_off:e_off + exp_counts[rank]])
        off += rc[s]
    # a numpy merge in source order (what k_merge_fold does), then the global gather
    keys = np.unique(recv[:, 0])
    sums = np.array([recv[recv[:, 0] == k, 2].sum() for k in keys])
    cnts = np.array([recv[recv[:, 0] == k, 4].sum() for k in keys])
    first = np.array([recv[recv[:, 0] == k, 5].min() for k in keys])
    got = D.gather_result({"key": keys, "sum": sums, "count": cnts}, first)
    # padded exchange: block [p] of my send buffer arrives as block [rank] ... of rank p's receive buffer
    cap = 8
    blocks = torch.zeros((world, cap + 1, W), dtype=torch.int64)
    for p_ in range(world):
        blocks[p_, 0, 0] = 3                                # header: record count
        blocks[p_, 1:4, 0] = torch.tensor([rank * 100 + p_ * 10 + j for j in range(3)])
    got_blocks = D.exchange_padded(blocks)
    for s in range(world):
        ok &= int(got_blocks[s, 0, 0]) == 3
        ok &= got_blocks[s, 1:4, 0].tolist() == [s * 100 + rank * 10 + j for j in range(3)]
    # resample anchor: start of the day of the GLOBAL first tick, also when a rank's shard is empty / before the epoch
    day = 86_400 * 10**9
    ok &= D.resample_anchor(5 * day + 123 + rank * day) == 5 * day
    ok &= D.resample_anchor(None if rank == 0 else 7 * day + 5) == 7 * day
    ok &= D.resample_anchor(-3 if rank == 1 else 10) == -day
    ok &= D.resample_anchor(None) == 0
    q.put((rank, ok, got["key"].tolist(), got["sum"].tolist(), got["count"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_exchange_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # expected global answer from all ranks' records
    allrec = np.concatenate([_fake_records(r, world, 300)[0] for r in range(world)])
    keys = np.unique(allrec[:, 0])
    first = np.array([allrec[allrec[:, 0] == k, 5].min() for k in keys])
    order = np.argsort(first, kind="stable")
    exp_keys = keys[order].tolist()
    exp_sum = [int(allrec[allrec[:, 0] == k, 2].sum()) for k in exp_keys]
    exp_cnt = [int(allrec[allrec[:, 0] == k, 4].sum()) for k in exp_keys]
    for rank, ok, k, s, c in results:
        assert ok, f"rank {rank}: exchange invariants"
        assert k == exp_keys and s == exp_sum and c == exp_cnt, f"rank {rank}: gathered result"


def test_shard_rows_and_owner():
    from pandasarrow_b200 import distributed as D
    n, world = 1_000_003, 8
    spans = [D.shard_rows(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    keys = np.arange(100_000, dtype=np.int64)
    own = D.owner_of(keys, world)
    assert own.min() == 0 and own.max() == world - 1
    assert abs(np.bincount(own).max() / np.bincount(own).min() - 1) < 0.1      # balanced partition
    assert (D.owner_of(keys, world, is_null=np.ones(len(keys), bool)) == 0).all()
