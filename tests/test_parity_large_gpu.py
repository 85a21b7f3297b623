"""Parity against the ORACLE at the sizes BASELINE.json's configs name, not only at test-sized
inputs: config 1 at 10 M x 1 K (materialising oracle = the reference constructor), config 2 at 100 M rows for
G in {16, 256, 4 K, 64 K, 1 M} and 20 M rows at G = 16 M with the radix-partition path forced, config 3 at 100 M
(int32 x dictionary key, nullable values), config 4 at 100 M ticks.  These reach the code that only large inputs
reach: 24-bit per-warp count words, 32-bit row numbers far above 2^24, the partition pre-pass, table growth.

Inputs come from the device generator (identical to hostgen.py, checked in test_groupby_gpu), are copied to the
host once for the oracle, and the two results are compared PER KEY with vectorised numpy (bit-exact for keys,
counts, min / max, first / last; <= 1e-12 relative for fp64 sum / mean — the bar BASELINE.json states).
Needs a GPU and ~8 GB of host RAM: -m gpu."""
import os

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

FP_RTOL = 1e-12
THREADS = min(os.cpu_count() or 1, 32)


@pytest.fixture(scope="module")
def env():
    import torch
    import pandasarrow_b200 as pab
    from oracle import oracle as orc
    return pab, orc, torch


def _np(a: pa.Array):
    return a.to_numpy(zero_copy_only=False)


def _cmp_vec(ours_keys, res, ora, column, aggs, what, rel_to_abs=None):
    """Per-key comparison: both results sorted by key.  NaN-free, null-free value columns unless masks say otherwise."""
    theirs = _np(ora.unique())
    assert len(ours_keys) == len(theirs), f"{what}: {len(ours_keys)} groups, oracle {len(theirs)}"
    so, st = np.argsort(ours_keys, kind="stable"), np.argsort(theirs, kind="stable")
    assert np.array_equal(ours_keys[so], theirs[st]), f"{what}: key sets differ"
    for a in aggs:
        if a == "mean":
            want_a, valid = ora.agg("mean", column, nthreads=THREADS, with_validity=True)
            wv = np.asarray(_np(valid), dtype=bool)[st]
        else:
            want_a = ora.agg(a, column, nthreads=THREADS)
            wv = np.asarray(want_a.is_valid())[st]
        got_a = res[a]
        assert got_a.type == want_a.type, f"{what} {a}: dtype {got_a.type} != {want_a.type}"
        gv = np.asarray(got_a.is_valid())[so]
        assert np.array_equal(gv, wv), f"{what} {a}: validity differs"
        g, w = _np(got_a)[so][gv], _np(want_a)[st][wv]
        if a in ("sum", "mean") and pa.types.is_floating(got_a.type):
            g, w = g.astype(np.float64), w.astype(np.float64)
            denom = np.maximum(np.abs(w), np.finfo(np.float64).tiny)
            rel = np.abs(g - w) / denom
            assert rel.max() <= FP_RTOL, f"{what} {a}: max rel err {rel.max():.3e}"
        else:
            assert np.array_equal(g, w), f"{what} {a}: {int((g != w).sum())} of {len(g)} groups differ"


def _first_appearance(keys_np):
    import pandas as pd
    return pd.unique(keys_np)


def _gen(pab, torch, n, G, scattered=False):
    k = torch.empty(n, dtype=torch.int64, device="cuda")
    v = torch.empty(n, dtype=torch.float64, device="cuda")
    pab.synth.keys(k, G)
    pab.synth.vals(v)
    if scattered:   # odd multiplier + offset: a bijection of int64, the key set is no longer a dense window
        k = k * 0x2545F4914F6CDD1D + 0x1234567
    torch.cuda.synchronize()
    return k, v


def test_config1_10m_rows_1k_groups_materialising_oracle(env):
    pab, orc, torch = env
    n, G = 10_000_000, 1000
    k, v = _gen(pab, torch, n, G)
    kh, vh = k.cpu().numpy(), v.cpu().numpy()
    rb = pa.record_batch({"k": pa.array(kh), "v": pa.array(vh)})
    ora = orc.OracleGroupBy(rb, "k", index=pa.array(np.arange(n, dtype=np.int64)), materialize=True)   # the reference ctor
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    aggs = ["sum", "mean", "count", "min", "max", "first", "last"]
    with pab.GroupBy("k", {"k": dk, "v": dv}) as gb:
        res = gb.aggregate(dv, aggs)
        assert gb.timing()["path"] == "lowcard"
        ours = _np(gb.unique())
    assert np.array_equal(ours, _first_appearance(kh))
    _cmp_vec(ours, res, ora, "v", aggs, "config 1")
    # and through host (Arrow) buffers, as a reference user would pass them
    with pab.GroupBy("k", rb) as gb:
        res_h = gb.aggregate(rb.column("v"), ["sum", "mean", "count"])
        for a in ("sum", "mean", "count"):
            assert res_h[a].equals(res[a]), a          # the shared-memory path is deterministic: bit for bit
    ora.close()


@pytest.mark.parametrize("G,scattered", [(16, False), (256, False), (1000, True), (4096, False), (4096, True),
                                         (65536, False), (1 << 20, False)])
def test_config2_100m_rows(env, G, scattered):
    pab, orc, torch = env
    n = 100_000_000
    k, v = _gen(pab, torch, n, G, scattered)
    kh, vh = k.cpu().numpy(), v.cpu().numpy()
    rb = pa.record_batch({"k": pa.array(kh), "v": pa.array(vh)})
    ora = orc.OracleGroupBy(rb, "k")
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    aggs = ["sum", "min", "max", "count"]
    with pab.GroupBy("k", {"k": dk, "v": dv}) as gb:
        res = gb.aggregate(dv, aggs)
        ours = _np(gb.unique())
        t = gb.timing()
    assert np.array_equal(ours, _first_appearance(kh)), "not in strict first-appearance order"
    _cmp_vec(ours, res, ora, "v", aggs, f"config 2 G={G} scattered={scattered} path={t['path']}/{t['mode']}")
    if G <= 1000:   # the narrow aggregate set takes another kernel variant: sum / mean / count
        with pab.GroupBy("k", {"k": dk, "v": dv}) as gb:
            res = gb.aggregate(dv, ["sum", "mean", "count"])
            ours = _np(gb.unique())
        _cmp_vec(ours, res, ora, "v", ["sum", "mean", "count"], f"config 2 narrow G={G} scattered={scattered}")
    ora.close()


def test_config2_16m_groups_partition_path_vs_oracle(env):
    pab, orc, torch = env
    n, G = 20_000_000, 1 << 24
    k, v = _gen(pab, torch, n, G)
    kh, vh = k.cpu().numpy(), v.cpu().numpy()
    rb = pa.record_batch({"k": pa.array(kh), "v": pa.array(vh)})
    ora = orc.OracleGroupBy(rb, "k")
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    aggs = ["sum", "min", "max", "count"]
    with pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G) as gb:
        res = gb.aggregate(dv, aggs)
        ours = _np(gb.unique())
        t = gb.timing()
    assert t["mode"] == "bucketed", t
    assert np.array_equal(ours, _first_appearance(kh))
    _cmp_vec(ours, res, ora, "v", aggs, "config 2 G=16M bucketed")
    ora.close()


@pytest.mark.parametrize("n,G,scattered,hint,ints,aggs", [
    (6_000_000, 50_000, True, False, False, ["sum", "mean", "count"]),                                  # unknown G: sketch decides, one level
    (6_000_000, 50_000, False, True, False, ["sum", "mean", "count", "min", "max", "first", "last"]),   # wide table
    (8_000_000, 3_000_000, True, True, False, ["sum", "min", "max", "count"]),                          # two levels
    (8_000_000, 3_000_000, False, False, True, ["sum", "mean", "count", "min", "max", "first", "last"]),   # two levels, integer values (double sum)
    (5_000_000, 5_000_000, True, False, False, ["sum", "count", "first"]),                              # every key once
])
def test_bucketed_path_vs_oracle(env, n, G, scattered, hint, ints, aggs):
    """bucketed.cuh (radix partition into buckets + shared-memory aggregation per bucket) against the oracle: one and
    two partition levels, narrow and wide tables, hinted and sketch-estimated group counts, the sentinel key value."""
    pab, orc, torch = env
    k, v = _gen(pab, torch, n, G, scattered)
    k[12345] = -7046029254386353131            # == kEmptyKey (0x9E3779B97F4A7C15): the table sentinel is a legal key
    k[n - 1] = -7046029254386353131
    if ints:
        v = ((v * 2001.0).to(torch.int64) - 1000)
    torch.cuda.synchronize()
    kh, vh = k.cpu().numpy(), v.cpu().numpy()
    rb = pa.record_batch({"k": pa.array(kh), "v": pa.array(vh)})
    ora = orc.OracleGroupBy(rb, "k")
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    with pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G if hint else 0) as gb:
        res = gb.aggregate(dv, aggs)
        ours = _np(gb.unique())
        t = gb.timing()
        assert t["mode"] == "bucketed", t
        ids = _np(gb.row_ids())
        assert np.array_equal(ours[ids[:100000]], kh[:100000])
    assert np.array_equal(ours, _first_appearance(kh)), "not in strict first-appearance order"
    if ints:
        for a in aggs:   # integer sums wrap exactly, means accumulate in double (1e-9 covers mixed signs, as in config 3)
            want = ora.agg(a, "v", nthreads=THREADS)
            so, st = np.argsort(ours, kind="stable"), np.argsort(_np(ora.unique()), kind="stable")
            g_, w_ = _np(res[a])[so], _np(want)[st]
            if a == "mean":
                assert (np.abs(g_ - w_) <= 1e-9 * np.maximum(np.abs(w_), 1.0)).all()
            else:
                assert np.array_equal(g_, w_), a
    else:
        _cmp_vec(ours, res, ora, "v", aggs, f"bucketed n={n} G={G} scattered={scattered} hint={hint} bits={t['replication']}")
    # a keys-only pass (group discovery without a value column) takes the same path
    with pab.GroupBy("k", {"k": dk}, expected_groups=G if hint else 0) as gb2:
        assert gb2.groupSize() == len(ours)
        assert np.array_equal(_np(gb2.unique()), ours)
    ora.close()


@pytest.mark.parametrize("n,G,ints", [(6_000_000, 50_000, False), (8_000_000, 3_000_000, False), (8_000_000, 3_000_000, True)])
def test_bucketed_path_nullable_values_vs_oracle(env, n, G, ints):
    """A nullable VALUE column on the bucketed path: the validity bit travels in bit 31 of the partitioned row numbers.
    One level (ranged aggregation into the global table) and two levels (whole buckets, records); with ~2.7 rows per
    group at G = 3 M many groups hold nothing but nulls (sum / min / max / mean null, count 0, first / last positional)."""
    pab, orc, torch = env
    k, v = _gen(pab, torch, n, G, scattered=True)
    if ints:
        v = ((v * 2001.0).to(torch.int64) - 1000)
    bits = torch.empty((n + 7) // 8, dtype=torch.uint8, device="cuda"); pab.synth.validity(bits, n, 0, 77, 3)   # every third value null
    torch.cuda.synchronize()
    kh = k.cpu().numpy()
    vmask = pa.py_buffer(bits.cpu().numpy())
    hv = pa.Array.from_buffers(pa.int64() if ints else pa.float64(), n, [vmask, pa.py_buffer(v.cpu().numpy())])
    rb = pa.record_batch({"k": pa.array(kh), "v": hv})
    ora = orc.OracleGroupBy(rb, "k")
    dk = pab.DeviceColumn.from_torch(k)
    dv = pab.DeviceColumn.from_torch(v, valid=bits, null_count=-1)
    aggs = ["sum", "mean", "count", "min", "max", "first", "last"]
    with pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G) as gb:
        res = gb.aggregate(dv, aggs)
        ours = _np(gb.unique())
        t = gb.timing()
        assert t["mode"] == "bucketed", t
    assert np.array_equal(ours, _first_appearance(kh)), "not in strict first-appearance order"
    theirs = _np(ora.unique())
    so, st = np.argsort(ours, kind="stable"), np.argsort(theirs, kind="stable")
    assert np.array_equal(ours[so], theirs[st])
    n_all_null = 0
    for a in aggs:
        if a == "mean":
            want, valid = ora.agg("mean", "v", nthreads=THREADS, with_validity=True)
            wv = np.asarray(_np(valid), dtype=bool)[st]
        else:
            want = ora.agg(a, "v", nthreads=THREADS)
            wv = np.asarray(want.is_valid())[st]
        got = res[a]
        assert got.type == want.type, a
        gv = np.asarray(got.is_valid())[so]
        assert np.array_equal(gv, wv), a
        if a == "sum":
            n_all_null = int((~wv).sum())
        g, w = _np(got)[so][gv], _np(want)[st][wv]
        if a in ("sum", "mean") and pa.types.is_floating(got.type):
            assert (np.abs(g - w) <= (FP_RTOL * np.maximum(np.abs(w), 1e-300) if not ints else 1e-9 * np.maximum(np.abs(w), 1.0))).all(), a
        else:
            assert np.array_equal(g, w), a
    if G >= 1_000_000:
        assert n_all_null > 1000          # the all-null-group case really occurred
    ora.close()


def test_config3_100m_rows_multi_key_nullable(env):
    pab, orc, torch = env
    n = 100_000_000
    k1 = torch.empty(n, dtype=torch.int64, device="cuda"); pab.synth.keys(k1, 1000, 0)
    k2 = torch.empty(n, dtype=torch.int64, device="cuda"); pab.synth.keys(k2, 64, 12345)
    k1i, k2i = k1.to(torch.int32), k2.to(torch.int32)
    del k1, k2
    v = torch.empty(n, dtype=torch.float64, device="cuda"); pab.synth.vals(v)
    iv = (v * 2001.0).to(torch.int64) - 1000
    bits = torch.empty((n + 7) // 8, dtype=torch.uint8, device="cuda"); pab.synth.validity(bits, n)
    kbits = torch.empty((n + 7) // 8, dtype=torch.uint8, device="cuda"); pab.synth.validity(kbits, n, 0, 31, 100)   # 1 % null keys
    torch.cuda.synchronize()
    words = pa.array([f"sym{i:02d}" for i in range(64)])
    vmask = pa.py_buffer(bits.cpu().numpy())
    kmask = pa.py_buffer(kbits.cpu().numpy())
    h_k1 = pa.Array.from_buffers(pa.int32(), n, [kmask, pa.py_buffer(k1i.cpu().numpy())])
    h_k2 = pa.DictionaryArray.from_arrays(pa.Array.from_buffers(pa.int32(), n, [None, pa.py_buffer(k2i.cpu().numpy())]), words)
    h_f = pa.Array.from_buffers(pa.float64(), n, [vmask, pa.py_buffer(v.cpu().numpy())])
    h_i = pa.Array.from_buffers(pa.int64(), n, [vmask, pa.py_buffer(iv.cpu().numpy())])
    rb = pa.record_batch({"k1": h_k1, "k2": h_k2, "f": h_f, "i": h_i})
    ora = orc.OracleGroupBy(rb, ["k1", "k2"])
    cf = pab.DeviceColumn.from_torch(v, valid=bits, null_count=-1)
    ci = pab.DeviceColumn.from_torch(iv, valid=bits, null_count=-1)
    aggs = ["sum", "mean", "count", "min", "max", "first", "last"]
    # keys as host Arrow arrays (the dictionary column tells the library its 6-bit index range: nullable int32 +
    # dictionary<64> pack into 40 bits), values device resident
    with pab.GroupBy(["k1", "k2"], {"k1": h_k1, "k2": h_k2, "f": cf, "i": ci}, expected_groups=65000) as gb:
        assert gb.groupSize() == ora.num_groups
        # composite key as one comparable number: (k1 or -1 for null) * 64 + k2
        def combo(u1, u2):
            a = np.where(np.asarray(u1.is_valid()), _np(u1.fill_null(0)).astype(np.int64), -1)
            if isinstance(u2, pa.DictionaryArray):
                u2 = u2.indices
            return a * 64 + _np(u2).astype(np.int64)
        ours = combo(gb.unique(0), gb.unique(1))
        theirs = combo(ora.unique(0), ora.unique(1))
        so, st = np.argsort(ours, kind="stable"), np.argsort(theirs, kind="stable")
        assert np.array_equal(ours[so], theirs[st])
        for col, name in ((cf, "f"), (ci, "i")):
            res = gb.aggregate(col, aggs)
            for a in aggs:
                if a == "mean":
                    want, valid = ora.agg("mean", name, nthreads=THREADS, with_validity=True)
                    wv = np.asarray(_np(valid), dtype=bool)[st]
                else:
                    want = ora.agg(a, name, nthreads=THREADS)
                    wv = np.asarray(want.is_valid())[st]
                got = res[a]
                assert got.type == want.type, (name, a)
                gv = np.asarray(got.is_valid())[so]
                assert np.array_equal(gv, wv), (name, a)
                g, w = _np(got)[so][gv], _np(want)[st][wv]
                if a in ("sum", "mean") and pa.types.is_floating(got.type):
                    rel = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
                    # integer means: sum(|x|) is the honest scale (values of both signs); 2000x headroom covers it
                    assert rel.max() <= (FP_RTOL if name == "f" else 1e-9), (name, a, rel.max())
                else:
                    assert np.array_equal(g, w), (name, a)
    ora.close()


def test_config4_100m_ticks_ohlc_sum(env):
    pab, orc, torch = env
    n = 100_000_000
    ts = torch.empty(n, dtype=torch.int64, device="cuda"); pab.synth.timestamps(ts)      # ~1000 ticks per minute
    v = torch.empty(n, dtype=torch.float64, device="cuda"); pab.synth.vals(v)
    torch.cuda.synchronize()
    idx = pa.Array.from_buffers(pa.timestamp("ns"), n, [None, pa.py_buffer(ts.cpu().numpy())])
    rb = pa.record_batch({"px": pa.array(v.cpu().numpy())})
    ora = orc.resample(rb, idx, 60 * 10**9)
    dts = pab.DeviceColumn.from_torch(ts, fmt="tsn:")
    dv = pab.DeviceColumn.from_torch(v)
    aggs = ["first", "max", "min", "last", "sum"]
    r = pab.resample({"px": dv}, dts, 60 * 10**9)
    res = r.aggregate(dv, aggs)
    assert r.timing()["path"] == "resample"
    ours = _np(r.index().cast(pa.int64()))
    assert (np.diff(ours) > 0).all()
    theirs = _np(ora.unique().cast(pa.int64()))
    st = np.argsort(theirs, kind="stable")
    assert np.array_equal(ours, theirs[st])
    for a in aggs:
        want = _np(ora.agg(a, "px", nthreads=THREADS))[st]
        got = _np(res[a])
        if a == "sum":
            assert (np.abs(got - want) / np.abs(want)).max() <= FP_RTOL
        else:
            assert np.array_equal(got, want), a
    r.close()
    ora.close()
