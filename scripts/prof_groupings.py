"""Device group materialisation timing: python scripts/prof_groupings.py --rows 200000000 --groups 1000"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pandasarrow_b200 as pab

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=200_000_000)
ap.add_argument("--groups", type=int, default=1000)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
n = a.rows
k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.keys(k, a.groups); pab.synth.vals(v)
torch.cuda.synchronize()
dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
for i in range(a.iters):
    g = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=a.groups)
    g.groupSize()
    g.groupings(rows=False)
    g.take_grouped(dv)
    print(f"iter {i}: groups {g.groupSize()} {g.groupings_timing()}")
    g.close()
