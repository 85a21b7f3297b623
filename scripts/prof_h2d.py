"""Staged (pageable) host-to-device ingest: GB/s of pa_column_to_device on an ordinary numpy-backed Arrow column, against the
pinned one-memcpy path.  PA_H2D_THREADS / PA_H2D_CHUNK_MB are read once per process: run one process per setting."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, torch
import pandasarrow_b200 as pab
ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=4096)
ap.add_argument("--iters", type=int, default=4)
a = ap.parse_args()
n = a.mb * (1 << 20) // 8
host = np.empty(n, dtype=np.int64); host[:] = 7          # touched, pageable
col = pa.array(host)                                      # zero-copy view
best = 0.0
for i in range(a.iters):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d = pab.to_device(col)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if i: best = max(best, 8.0 * n / dt / 1e9)
    del d
print(f"threads={os.environ.get('PA_H2D_THREADS','default')} chunk_mb={os.environ.get("PA_H2D_CHUNK_MB","8")} bufs={os.environ.get("PA_H2D_BUFS","2")} pageable {best:.1f} GB/s")
