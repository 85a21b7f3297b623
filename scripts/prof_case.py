"""One small device-resident group-by, for ncu / timing experiments.
python scripts/prof_case.py --rows 100000000 --groups 1000 --aggs sum,mean,count --iters 3 [--path global]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pandasarrow_b200 as pab

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--groups", type=int, default=1000)
ap.add_argument("--aggs", default="sum,mean,count")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--path", default="auto")
ap.add_argument("--hint", type=int, default=0)
ap.add_argument("--no-partition", action="store_true")
ap.add_argument("--hashed", action="store_true", help="scramble the dense key ids into random-looking 64-bit keys")
a = ap.parse_args()
n = a.rows
k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.keys(k, a.groups); pab.synth.vals(v)
if a.hashed:   # bijective scramble: multiply, xor with a logical right shift, multiply
    k.mul_(-3335678366873096957)
    k.bitwise_xor_((k >> 29) & ((1 << 35) - 1))
    k.mul_(-4658895280553007687)
torch.cuda.synchronize()
dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
g = pab.GroupBy("k", {"k": dk, "v": dv}, path=a.path, expected_groups=a.hint, no_partition=a.no_partition)
for i in range(a.iters):
    g.aggregate(dv, a.aggs.split(","), fetch=False)
    t = g.timing()
    print(f"iter {i}: {t} rows/s={n / t['total_ms'] * 1e3:.3e} GB/s(scan)={16 * n / t['scan_ms'] / 1e6:.1f}")
print("groups", g.groupSize())
