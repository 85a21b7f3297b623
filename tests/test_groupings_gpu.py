"""Group materialisation on the device: pa_groupby_groupings / pa_groupby_take_grouped against
what the reference computes with Grouper::MakeGroupings / ApplyGroupings (dataframe.cpp:1546,1562,1586-1588) —
restated with numpy from the group ids, and through the oracle's own ApplyGroupings slices.  Needs a GPU: -m gpu."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


def _expected(ids, G):
    rows = np.argsort(ids, kind="stable").astype(np.int32)
    offsets = np.concatenate([[0], np.cumsum(np.bincount(ids, minlength=G))]).astype(np.int32)
    return offsets, rows


def _frame(n, G, seed):
    rng = np.random.default_rng(seed)
    vm = rng.random(n) < 0.15
    return pa.record_batch({
        "k": pa.array(rng.integers(0, G, n) * 31 + 5, pa.int64(), mask=rng.random(n) < 0.03),
        "f": pa.array(rng.normal(size=n), pa.float64(), mask=vm),
        "i32": pa.array(rng.integers(-1000, 1000, n).astype(np.int32), pa.int32()),
        "u8": pa.array(rng.integers(0, 255, n).astype(np.uint8), pa.uint8(), mask=vm),
        "s16": pa.array(rng.integers(-300, 300, n).astype(np.int16), pa.int16()),
        "f32": pa.array(rng.normal(size=n).astype(np.float32), pa.float32(), mask=vm),
        "ts": pa.array(rng.integers(0, 10**15, n), pa.timestamp("ns"))})


@pytest.mark.parametrize("n,G,kw", [(1, 1, {}), (1000, 1, {}), (100_003, 7, {}), (200_000, 1000, {}), (300_001, 5000, {}),
                                    (400_000, 120_000, {"expected_groups": 120_000}), (250_000, 250_000, {"path": "global"})])
def test_groupings_and_take(pab, n, G, kw):
    from oracle import oracle as orc
    rb = _frame(n, G, seed=n + G)
    gb = pab.GroupBy("k", rb, **kw)
    ids = gb.row_ids().to_numpy()
    ng = gb.groupSize()
    offsets, rows = gb.groupings()
    want_off, want_rows = _expected(ids, ng)
    assert offsets.type == pa.int32() and rows.type == pa.int32()
    assert np.array_equal(offsets.to_numpy(), want_off)
    assert np.array_equal(rows.to_numpy(), want_rows)                  # ascending inside every group (stable)
    only_off, none = gb.groupings(rows=False)
    assert none is None and only_off.equals(offsets)
    take = pa.array(want_rows)
    for c in ("f", "i32", "u8", "s16", "f32", "ts"):
        got = gb.take_grouped(rb.column(c))
        assert got.type == rb.column(c).type
        assert got.equals(rb.column(c).take(take)), c
    # the oracle's ApplyGroupings slices (group j of the reference) for a few groups
    # (matched by key: Grouper's id order may swap keys that first appear in one mini-batch, see tests/util.py)
    from util import align_to
    ora = orc.OracleGroupBy(rb, "k")
    perm = align_to([(x,) for x in gb.unique().to_pylist()], [(x,) for x in ora.unique().to_pylist()])
    got = gb.take_grouped(rb.column("f"))
    o = offsets.to_numpy()
    for j in sorted({0, ng // 2, ng - 1}):
        m = perm[j]
        assert got.slice(o[m], o[m + 1] - o[m]).equals(ora.group_slice("f", j)), j
    t = gb.groupings_timing()
    assert t["build_ms"] > 0 and t["take_ms"] >= 0


def test_groupings_device_columns_sliced_input_and_resampler(pab):
    import torch
    rng = np.random.default_rng(5)
    n = 50_000
    k = torch.from_numpy(rng.integers(0, 300, n)).cuda()
    v = torch.from_numpy(rng.normal(size=n)).cuda()
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    gb = pab.GroupBy("k", {"k": dk, "v": dv})
    offsets, rows = gb.groupings()
    want_off, want_rows = _expected(gb.row_ids().to_numpy(), gb.groupSize())
    assert np.array_equal(rows.to_numpy(), want_rows) and np.array_equal(offsets.to_numpy(), want_off)
    assert np.array_equal(gb.take_grouped(dv).to_numpy(), v.cpu().numpy()[want_rows])
    # host column with a non-zero offset and nulls
    col = pa.array(rng.normal(size=n + 13), mask=rng.random(n + 13) < 0.2).slice(13)
    assert gb.take_grouped(col).equals(col.take(pa.array(want_rows)))
    # resampler: buckets in time order, rows ascending -> the identity permutation on a sorted index
    ts = np.cumsum(rng.integers(1, 2_000_000_000, n)).astype(np.int64)
    rs = pab.resample({"v": pa.array(rng.normal(size=n))}, pa.array(ts, pa.timestamp("ns")), 60_000_000_000)
    o, r = rs.groupings()
    assert np.array_equal(r.to_numpy(), np.arange(n, dtype=np.int32))
    assert o.to_numpy()[0] == 0 and o.to_numpy()[-1] == n and len(o) == rs.groupSize() + 1


def test_groupings_empty_and_errors(pab):
    e = pa.record_batch({"k": pa.array([], pa.int64()), "v": pa.array([], pa.float64())})
    g = pab.GroupBy("k", e)
    o, r = g.groupings()
    assert o.to_pylist() == [0] and len(r) == 0
    assert len(g.take_grouped(e.column("v"))) == 0
    rb = pa.record_batch({"k": pa.array([1, 2, 1], pa.int64()), "v": pa.array([1.0, 2.0, 3.0])})
    g = pab.GroupBy("k", rb)
    with pytest.raises(pab.PaError, match="rows"):
        g.take_grouped(pa.array([1.0, 2.0]))
    assert g.take_grouped(rb.column("v")).to_pylist() == [1.0, 3.0, 2.0]
    assert g.take_grouped(pa.array([True, False, True])).to_pylist() == [True, True, False]   # bit-packed booleans: bit gather


@pytest.mark.parametrize("G", [3, 1000, 1500, 70_000, 1_200_000])
def test_counting_sort_passes_and_scatter_take(pab, G):
    """MakeGroupings = own stable counting sort (csort.cuh): one pass up to 1024 groups, two up to 2^20, three beyond;
    ApplyGroupings = scatter through dest[row].  Checked against a stable numpy argsort of the row ids."""
    rng = np.random.default_rng(G)
    n = 2_000_003
    k = rng.integers(0, G, n)
    k[:min(G, n)] = np.arange(min(G, n))
    rb = pa.record_batch({"k": pa.array(k, pa.int64())})
    gb = pab.GroupBy("k", rb, expected_groups=G)
    offsets, rows = gb.groupings()
    ids = gb.row_ids().to_numpy()
    want_off, want_rows = _expected(ids, gb.groupSize())
    assert np.array_equal(offsets.to_numpy(), want_off)
    assert np.array_equal(rows.to_numpy(), want_rows)
    for col in (pa.array(rng.normal(size=n), mask=rng.random(n) < 0.1), pa.array(rng.integers(-100, 100, n).astype(np.int16)),
                pa.array(rng.integers(0, 255, n).astype(np.uint8), mask=rng.random(n) < 0.5), pa.array(rng.random(n) < 0.3, mask=rng.random(n) < 0.05),
                pa.array(rng.normal(size=n + 5).astype(np.float32)).slice(5)):
        assert gb.take_grouped(col).equals(col.take(pa.array(want_rows))), col.type
