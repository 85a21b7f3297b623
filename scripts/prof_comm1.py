"""Single-GPU run of the sharded step (pa_comm of world size 1): phases + the merged handle's own stage times.
python scripts/prof_comm1.py [rows] [groups] [steps]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import pandasarrow_b200 as pab
from pandasarrow_b200 import distributed as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.cuda.set_device(0)
dist.init_process_group("gloo", store=dist.FileStore(os.path.join(tempfile.mkdtemp(), "s"), 1), rank=0, world_size=1)
k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.keys(k, G, 0); pab.synth.vals(v, 0); torch.cuda.synchronize()
dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
gb = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G)
comm = D.Comm(device=0)
aggs = ["sum", "mean", "count"]
for it in range(steps):
    m = comm.sharded_aggregate(gb, dv, aggs)
    t = m.timing()
    print(f"it{it}: phases {comm.phases()} info {comm.exchange_info()} merged: fill {t['pack_ms']:.2f} insert+compact {t['scan_ms']:.2f} sort+fold {t['merge_ms']:.2f} emit {t['emit_ms']:.2f} groups {m.groupSize()}")
    m.close()
comm.close(); gb.close()
dist.destroy_process_group()
