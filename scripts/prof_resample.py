"""Device-resident resample (config 4): sorted ticks, 1-minute buckets, OHLC + sum."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pandasarrow_b200 as pab
ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=400_000_000)
ap.add_argument("--aggs", default="first,max,min,last,sum")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
n = a.rows
ts = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
pab.synth.timestamps(ts); pab.synth.vals(v); torch.cuda.synchronize()
dts = pab.DeviceColumn.from_torch(ts, fmt="tsn:"); dv = pab.DeviceColumn.from_torch(v)
r = pab.resample({"v": dv}, dts, 60 * 10**9)
for i in range(a.iters):
    r.aggregate(dv, a.aggs.split(","), fetch=False)
    t = r.timing()
    print(f"iter {i}: {t} rows/s={n / t['total_ms'] * 1e3:.3e} GB/s(scan)={16 * n / t['scan_ms'] / 1e6:.1f}")
print("buckets", r.groupSize())
