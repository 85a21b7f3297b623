"""Bucketed path (bucketed.cuh), one configuration, for timing / ncu launch lists.
python scripts/prof_bucketed.py --rows 1000000000 --groups 1048576 [--aggs sum,min,max,count] [--scattered]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pandasarrow_b200 as pab

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000_000)
ap.add_argument("--groups", type=int, default=1 << 20)
ap.add_argument("--aggs", default="sum,min,max,count")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--scattered", action="store_true")
ap.add_argument("--no-hint", action="store_true")
ap.add_argument("--no-partition", action="store_true")
ap.add_argument("--bits", type=lambda x: int(x, 0), default=0)
a = ap.parse_args()
n = a.rows
k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.keys(k, a.groups); pab.synth.vals(v)
if a.scattered:
    k.mul_(0x2545F4914F6CDD1D).add_(0x1234567)
torch.cuda.synchronize()
dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
g = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=0 if a.no_hint else a.groups, no_partition=a.no_partition, bucket_bits=a.bits)
for i in range(a.iters):
    g.aggregate(dv, a.aggs.split(","), fetch=False)
    t = g.timing()
    print(f"iter {i}: {t} rows/s={n / t['total_ms'] * 1e3:.3e}")
print("groups", g.groupSize())
