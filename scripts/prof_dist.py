"""Phase timing of the multi-GPU step (torchrun --nproc-per-node N scripts/prof_dist.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import pandasarrow_b200 as pab
from pandasarrow_b200 import distributed as D
from pandasarrow_b200._lib import PA_PARTIAL_WORDS
from pandasarrow_b200.groupby import MergedGroupBy

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.keys(k, G, rank * n); pab.synth.vals(v, rank * n); torch.cuda.synchronize()
dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
gb = pab.GroupBy("k", {"k": dk, "v": dv}, row_base=rank * n, expected_groups=G if G > 1024 else 0)
aggs = ["sum", "mean", "count"]
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(6):
    dist.barrier(); t0 = T()
    gb.aggregate(dv, aggs, fetch=False); t1 = T()
    counts = gb.partials_count(world); t2 = T()
    send = torch.empty((max(sum(counts), 1), PA_PARTIAL_WORDS), dtype=torch.int64, device="cuda")
    gb.partials_export(world, send.data_ptr(), send.shape[0]); t3 = T()
    recv, rc = D.exchange_records(send[:sum(counts)], counts); t4 = T()
    m = MergedGroupBy(recv.data_ptr(), rc, aggs, "g", "l", device=local); t5 = T()
    m.close(); t6 = T()
    if G <= D.PADDED_BLOCK_RECORDS:
        # padded path, phase by phase
        cap = D.PADDED_BLOCK_RECORDS
        sendb = torch.empty((world, cap + 1, PA_PARTIAL_WORDS), dtype=torch.int64, device="cuda"); p0 = T()
        gb.partials_export_padded(world, sendb.data_ptr(), cap); p1 = T()
        recvb = D.exchange_padded(sendb); p2 = T()
        m2 = MergedGroupBy(recvb.data_ptr(), [0] * world, aggs, "g", "l", device=local, padded_block_records=cap); p3 = T()
        gs = m2.groupSize(); tm = gb.timing(); p4 = T()
        m2.close(); p5 = T()
        # the same without intermediate synchronisation
        q0 = T()
        sendb = torch.empty((world, cap + 1, PA_PARTIAL_WORDS), dtype=torch.int64, device="cuda")
        gb.partials_export_padded(world, sendb.data_ptr(), cap)
        recvb = D.exchange_padded(sendb)
        m2 = MergedGroupBy(recvb.data_ptr(), [0] * world, aggs, "g", "l", device=local, padded_block_records=cap)
        m2.close(); q1 = T()
        if rank == 0:
            print(f"it{it} padded: alloc {1e3*(p0-t6):.3f} export {1e3*(p1-p0):.3f} exchange {1e3*(p2-p1):.3f} merge {1e3*(p3-p2):.3f} query {1e3*(p4-p3):.3f} close {1e3*(p5-p4):.3f}; unsynced total {1e3*(q1-q0):.3f} ms; groups {gs}")
    if rank == 0:
        print(f"it{it}: local {1e3*(t1-t0):.3f}  count {1e3*(t2-t1):.3f}  export {1e3*(t3-t2):.3f}  exchange {1e3*(t4-t3):.3f}  merge {1e3*(t5-t4):.3f}  close {1e3*(t6-t5):.3f} ms; groups {m.groupSize() if False else sum(rc)}")
dist.destroy_process_group()
