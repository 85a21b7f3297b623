"""BASELINE.json's full size (1 B rows per GPU) through size-independent properties — the oracle cannot run there
(it needs ~4x the input in RAM and int32 row ids).  Device-generated inputs (the benchmark generator), checked with:
counts add up, group sums add up to the column total, mean * count = sum, exact linearity under a power-of-two
scale, run-to-run determinism of the shared-memory path, the closed-form key set, distinct first rows, time order
and OHLC ordering for resample, var(2x) = 4 var(x).  Needs a GPU with >= 40 GB free: -m gpu."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

N = 1_000_000_000


@pytest.fixture(scope="module")
def data():
    import torch
    import pandasarrow_b200 as pab
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2**30:
        pytest.skip("needs 60 GB of free device memory")
    v = torch.empty(N, dtype=torch.float64, device="cuda")
    pab.synth.vals(v)
    torch.cuda.synchronize()
    return pab, torch, v


def _keys(pab, torch, G):
    k = torch.empty(N, dtype=torch.int64, device="cuda")
    pab.synth.keys(k, G)
    torch.cuda.synchronize()
    return k


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


@pytest.mark.parametrize("G,path", [(1000, "lowcard"), (4096, "global"), (1 << 20, "global")])
def test_one_billion_rows_invariants(data, G, path):
    pab, torch, v = data
    k = _keys(pab, torch, G)
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    gb = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G if G > 1024 else 0)
    r = gb.aggregate(dv, ["sum", "mean", "count", "min", "max"])
    assert gb.timing()["path"] == path
    assert gb.groupSize() == G
    keys = gb.unique().to_numpy()
    assert np.array_equal(np.sort(keys), np.arange(G))                      # key[i] = splitmix64(i ^ seed) % G
    cnt, s, m = r["count"].to_numpy(), r["sum"].to_numpy(), r["mean"].to_numpy()
    assert cnt.sum() == N and cnt.min() > 0
    total = float(v.sum().item())
    assert _rel(float(np.sum(s)), total) <= 1e-12                           # checksum of checksums
    assert np.max(np.abs(m * cnt - s) / s) <= 1e-12
    mn, mx = r["min"].to_numpy(), r["max"].to_numpy()
    assert (mn <= m).all() and (m <= mx).all() and mn.min() == float(v.min().item()) and mx.max() == float(v.max().item())
    fr = gb.first_rows().to_numpy()
    assert len(np.unique(fr)) == G and (np.diff(fr.astype(np.int64)) > 0).all() and fr[0] == 0    # first-appearance order
    first_key = k[torch.from_numpy(fr[:1000].astype(np.int64)).cuda()].cpu().numpy()
    assert np.array_equal(first_key, keys[:1000])
    # linearity under an exact scale: sum(2^-3 x) = 2^-3 sum(x) bit for bit where the order of additions is fixed
    w = v * 0.125
    dw = pab.DeviceColumn.from_torch(w)
    r2 = gb.aggregate(dw, ["sum", "count"])
    assert np.array_equal(r2["count"].to_numpy(), cnt)
    if path == "lowcard":
        assert np.array_equal(r2["sum"].to_numpy(), s * 0.125)
        again = gb.aggregate(dv, ["sum", "mean", "count"])
        assert np.array_equal(again["sum"].to_numpy(), s) and np.array_equal(again["mean"].to_numpy(), m)   # determinism
    else:
        assert np.max(np.abs(r2["sum"].to_numpy() - s * 0.125) / (s * 0.125)) <= 1e-12
    # second-stage aggregates: var(2^-3 x) = 2^-6 var(x)
    va = gb.aggregate(dv, ["variance"])["variance"].to_numpy()
    vb = gb.aggregate(dw, ["variance"])["variance"].to_numpy()
    assert np.max(np.abs(vb - va / 64) / (va / 64)) <= 1e-12
    assert abs(va.mean() - 1.0 / 12) < 1e-3 and np.max(np.abs(va - 1.0 / 12)) < 3e-2     # values are uniform on [0, 1)
    del w, k


def test_one_billion_ticks_resample(data):
    pab, torch, v = data
    ts = torch.empty(N, dtype=torch.int64, device="cuda")
    pab.synth.timestamps(ts)
    torch.cuda.synchronize()
    dts, dv = pab.DeviceColumn.from_torch(ts, fmt="tsn:"), pab.DeviceColumn.from_torch(v)
    rs = pab.resample({"v": dv}, dts, 60 * 10**9)
    r = rs.aggregate(dv, ["first", "max", "min", "last", "sum", "count"])
    labels = rs.index().cast(pa.int64()).to_numpy()
    assert (np.diff(labels) > 0).all() and (labels % (60 * 10**9) == 0).all()
    assert labels[0] <= int(ts[0].item()) < labels[0] + 60 * 10**9 and labels[-1] <= int(ts[-1].item()) < labels[-1] + 60 * 10**9
    cnt = r["count"].to_numpy()
    assert cnt.sum() == N
    o, h, l, c = (r[a].to_numpy() for a in ("first", "max", "min", "last"))
    assert (l <= o).all() and (o <= h).all() and (l <= c).all() and (c <= h).all()
    assert _rel(float(np.sum(r["sum"].to_numpy())), float(v.sum().item())) <= 1e-12
    assert o[0] == float(v[0].item()) and c[-1] == float(v[-1].item())
    # bucket boundaries: count of the first bucket = ticks below the second label
    assert cnt[0] == int((ts < int(labels[1])).sum().item())
