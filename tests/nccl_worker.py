"""Worker of tests/test_nccl_multigpu.py: one process per GPU (torchrun), real NCCL.  Every rank aggregates its
row-range shard, the partials travel over NCCL by each of the three transports, every rank fetches the groups it
owns, rank 0 gathers them, restores global first-appearance order and compares with the ORACLE run on the whole
data set.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALL = ["sum", "mean", "count", "min", "max", "first", "last"]


def main():
    import faulthandler
    faulthandler.enable()
    import torch
    import torch.distributed as dist
    import pandasarrow_b200 as pab
    from pandasarrow_b200 import distributed as D
    from pandasarrow_b200 import hostgen as hg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    comm = D.Comm(device=local)
    n = 2_000_003                                   # rows per rank (odd: partial row groups)
    stream = torch.cuda.current_stream(dev).cuda_stream
    failures = []
    for G, scattered in ((1000, False), (1000, True), (60_000, False), (60_000, True)):
        first = rank * n
        k = torch.empty(n, dtype=torch.int64, device=dev); pab.synth.keys(k, G, first)
        v = torch.empty(n, dtype=torch.float64, device=dev); pab.synth.vals(v, first)
        if scattered:
            k = k * 0x2545F4914F6CDD1D + 0x1234567
        torch.cuda.synchronize()
        dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
        gb = pab.GroupBy("k", {"k": dk, "v": dv}, device=local, stream=stream, row_base=first)
        NARROW = ["sum", "mean", "count"]             # travels as 32-byte compact records (merge.cuh)
        transports = [("c-abi", lambda: comm.sharded_aggregate(gb, dv, ALL), ALL),
                      ("c-abi compact records", lambda: comm.sharded_aggregate(gb, dv, NARROW), NARROW),
                      ("counted", lambda: D.sharded_aggregate(gb, dv, ALL, "g", "l", stream=stream, padded=False), ALL)]
        if G <= D.PADDED_BLOCK_RECORDS:
            transports.append(("padded", lambda: D.sharded_aggregate(gb, dv, ALL, "g", "l", stream=stream, padded=True), ALL))
        if rank == 0:
            from oracle import oracle as orc
            kh = hg.keys(world * n, G)
            if scattered:
                with np.errstate(over="ignore"):
                    kh = (kh.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D) + np.uint64(0x1234567)).astype(np.int64)
            vh = hg.vals(world * n)
            rb = pa.record_batch({"k": pa.array(kh), "v": pa.array(vh)})
            ora = orc.OracleGroupBy(rb, "k")
            theirs = ora.unique().to_numpy()
            st = np.argsort(theirs, kind="stable")
            want = {a: ora.agg(a, "v", nthreads=8).to_numpy()[st] for a in ALL}
            import pandas as pd
            fa = pd.unique(kh)
        for name, run, aggs in transports:
            m = run()
            owned = {a: m.fetch(a).to_numpy(zero_copy_only=False) for a in aggs}
            owned["key"] = m.unique().to_numpy()
            fr = m.first_rows().to_numpy()
            own = D.owner_of(owned["key"], world)
            if not (own == rank).all():
                failures.append(f"{name} G={G}: rank {rank} merged a key it does not own")
            m.close()
            res = D.gather_result(owned, fr)
            if rank == 0:
                tag = f"{name} G={G} scattered={scattered}"
                if not np.array_equal(res["key"], fa):
                    failures.append(f"{tag}: keys / global first-appearance order differ")
                    continue
                so = np.argsort(res["key"], kind="stable")
                for a in aggs:
                    g_, w_ = res[a][so], want[a]
                    if a in ("sum", "mean"):
                        rel = np.abs(g_ - w_) / np.abs(w_)
                        if not rel.max() <= 1e-12:
                            failures.append(f"{tag} {a}: rel err {rel.max():.3e}")
                    elif not np.array_equal(g_, w_):
                        failures.append(f"{tag} {a}: {(g_ != w_).sum()} groups differ")
        gb.close()
        if rank == 0:
            ora.close()
    if rank == 0:
        print("phases of the last c-abi step:", comm.phases())
    bad = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(bad)
    for f in failures:
        print(f"[rank {rank}] FAIL {f}")
    comm.close()
    dist.destroy_process_group()
    if bad.item():
        sys.exit(1)
    if rank == 0:
        print("NCCL PARITY OK")


if __name__ == "__main__":
    main()
