import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import pandasarrow_b200 as pab
from pandasarrow_b200 import hostgen as hg

def first_appearance(k):
    _, idx = np.unique(k, return_index=True)
    order = np.sort(idx)
    return k[order], order

for n, G in [(300000, 256), (300000, 16), (100000, 256), (3000, 256), (1_000_000, 1000)]:
    k = hg.keys(n, G); v = hg.vals(n)
    rb = pa.record_batch({"k": pa.array(k), "v": pa.array(v)})
    want, rows = first_appearance(k)
    for path in ["auto", "global"]:
        g = pab.GroupBy("k", rb, path=path)
        ng = g.groupSize()
        u0 = g.unique().to_numpy()
        r = g.aggregate(rb.column("v"), ["sum", "mean", "count"])
        u = g.unique().to_numpy()
        t = g.timing()
        bad = np.nonzero(u[:len(want)] != want[:len(u)])[0]
        bad0 = np.nonzero(u0[:len(want)] != want[:len(u0)])[0]
        cnt = r["count"].to_numpy()
        wantcnt = np.array([np.sum(k == x) for x in want[:50]])
        print(f"n={n} G={G} path={path}->{t['path']} ng={ng} len(u)={len(u)} keysonly_bad={len(bad0)} bad={len(bad)} first_bad={bad[:5]} "
              f"u[bad]={u[bad[:5]]} want[bad]={want[bad[:5]]} rows[bad]={rows[bad[:5]]} sumcnt={cnt.sum()} cnt_ok={np.array_equal(cnt[:50], wantcnt)}")
