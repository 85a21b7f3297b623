"""Second-stage aggregates — product, variance, stddev (reference:
GROUPBY_AGG(product), GROUPBY_NUMERIC_AGG(variance|stddev), dataframe.cpp:1516-1536) — against the
oracle's per-group arrow::compute calls.  Needs a GPU: -m gpu.

Bars: integer products bit-exact (wrapping); variance / stddev within 1e-12 relative to
max(value, mean(x^2)) resp. max(value, sqrt(mean(x^2))) of the group (a constant group's variance is
rounding noise around 0 in both implementations); float products within 1e-12 relative while the
product stays a normal number (the test data keep it near 1)."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

S2 = ["product", "variance", "stddev"]


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _frame(n, G, seed, null_keys=True, null_vals=True):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, G, n) * 7919 - 3           # not dense
    sign = np.where(rng.random(n) < 0.3, -1.0, 1.0)
    f = sign * np.exp(rng.uniform(-0.01, 0.01, n))  # products stay near +-1
    w = rng.normal(3.0, 2.0, n)                     # variance column
    i = rng.integers(-5, 6, n)
    i[i == 0] = 3                                   # (a zero factor would make every product 0)
    vm = rng.random(n) < 0.1 if null_vals else np.zeros(n, bool)
    km = rng.random(n) < 0.02 if null_keys else np.zeros(n, bool)
    return pa.record_batch({
        "k": pa.array(k, pa.int64(), mask=km),
        "f": pa.array(f, pa.float64(), mask=vm), "w": pa.array(w, pa.float64(), mask=vm),
        "w32": pa.array(w.astype(np.float32), pa.float32(), mask=vm),
        "i": pa.array(i, pa.int64(), mask=vm), "i32": pa.array(i.astype(np.int32), pa.int32(), mask=vm),
        "u16": pa.array((i + 6).astype(np.uint16), pa.uint16(), mask=vm)})


def _numeric(ora, func, col):
    """GROUPBY_NUMERIC_AGG drops validity; the oracle reports it separately and so does the C ABI."""
    want, valid = ora.agg(func, col, nthreads=8, with_validity=True)
    v = np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool)
    return pa.array(want.to_numpy(zero_copy_only=False), pa.float64(), mask=~v)


def _second_moment(rb, ora, col):
    """mean(x^2) per group, oracle order (NaN/inf-free)."""
    import pyarrow.compute as pc
    from oracle import oracle as orc
    x = rb.column(col).cast(pa.float64())
    o2 = orc.OracleGroupBy(pa.record_batch({"k": rb.column("k"), "sq": pc.multiply(x, x)}), "k")
    m = o2.agg("mean", "sq", nthreads=8).to_numpy(zero_copy_only=False).astype(np.float64)
    return np.where(np.isfinite(m), m, 0.0)


def _check(pab, orc, rb, what, cols_var=("w", "w32", "i", "i32", "u16"), cols_prod=("f", "i", "i32", "u16"), **kw):
    from util import align_to, assert_exact, assert_fp_close, first_appearance_order
    gb = pab.GroupBy("k", rb, **kw)
    ora = orc.OracleGroupBy(rb, "k")
    ours = [(x,) for x in gb.unique().to_pylist()]
    assert ours == first_appearance_order([rb.column("k")])
    perm = pa.array(align_to(ours, [(x,) for x in ora.unique().to_pylist()]))
    for c in cols_var:
        r = gb.aggregate(rb.column(c), ["variance", "stddev", "mean"])
        m2 = _second_moment(rb, ora, c)
        assert_fp_close(r["variance"].take(perm), _numeric(ora, "variance", c), f"{what} variance({c})", m2)
        assert_fp_close(r["stddev"].take(perm), _numeric(ora, "stddev", c), f"{what} stddev({c})", np.sqrt(m2))
        assert_fp_close(r["mean"].take(perm), _numeric(ora, "mean", c), f"{what} mean({c})", np.sqrt(m2))
    for c in cols_prod:
        got = gb.product(c).take(perm)
        want = ora.agg("product", c, nthreads=8)
        if pa.types.is_floating(rb.column(c).type):
            assert_fp_close(got, want, f"{what} product({c})")
        else:
            assert_exact(got, want, f"{what} product({c})")
    gb.close()


@pytest.mark.parametrize("n,G,kw", [
    (60_000, 5, {}),                                   # warp-combined updates
    (60_000, 64, {}),
    (200_000, 300, {}),                                # CTA-shared accumulators, shared-memory first pass
    (200_000, 3000, {}),                               # CTA-shared accumulators, global-table first pass
    (200_000, 4096, {"path": "global"}),
    (300_000, 20_000, {}),                             # L2 atomics
    (300_000, 150_000, {"expected_groups": 150_000}),
])
def test_stage2_matches_oracle(pab, orc, n, G, kw):
    _check(pab, orc, _frame(n, G, seed=G), f"n={n} G={G}", **kw)


def test_stage2_no_nulls_and_device_columns(pab, orc):
    import torch
    rb = _frame(100_000, 50, seed=9, null_keys=False, null_vals=False)
    _check(pab, orc, rb, "no nulls")
    # device-resident inputs (no validity): same numbers as host inputs
    k = torch.from_numpy(rb.column("k").to_numpy()).cuda()
    w = torch.from_numpy(rb.column("w").to_numpy()).cuda()
    dk, dw = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(w)
    gd = pab.GroupBy("k", {"k": dk, "w": dw})
    gh = pab.GroupBy("k", rb)
    a, b = gd.aggregate(dw, ["variance", "count"]), gh.aggregate(rb.column("w"), ["variance", "count"])
    assert a["count"].equals(b["count"])
    np.testing.assert_allclose(a["variance"].to_numpy(), b["variance"].to_numpy(), rtol=1e-13)


def test_stage2_edge_cases(pab, orc):
    # all-null group, single-row group, constant group, NaN, +-inf; outputs requested together with first-pass ones
    k = pa.array([1, 1, 2, 2, 2, 3, 4, 4, 4, 5, 5, None, None], pa.int64())
    v = pa.array([None, None, 2.5, 2.5, 2.5, 7.0, 1.0, float("nan"), 3.0, float("inf"), 1.0, 4.0, 6.0], pa.float64())
    rb = pa.record_batch({"k": k, "v": v})
    gb = pab.GroupBy("k", rb)
    ora = orc.OracleGroupBy(rb, "k")
    r = gb.aggregate(v, ["sum", "count", "product", "variance", "stddev"])
    assert gb.unique().to_pylist() == ora.unique().to_pylist() == [1, 2, 3, 4, 5, None]
    assert r["count"].to_pylist() == [0, 3, 1, 3, 2, 2]
    for a in ("variance", "stddev"):
        want = _numeric(ora, a, "v")
        got = r[a]
        assert got.is_valid().to_pylist() == want.is_valid().to_pylist() == [False, True, True, True, True, True]
        g, w = got.to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False)
        assert np.isnan(g[3]) and np.isnan(w[3]) and np.isnan(g[4]) and np.isnan(w[4])     # NaN member; inf - inf
        assert g[1] == w[1] == 0.0 and g[2] == w[2] == 0.0
        assert g[5] == w[5]
    want = ora.agg("product", "v")
    assert r["product"].is_valid().to_pylist() == want.is_valid().to_pylist()
    g, w = r["product"].to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False)
    assert g[1] == w[1] == 15.625 and g[2] == 7.0 and np.isnan(g[3]) and np.isnan(w[3]) and g[4] == w[4] == float("inf") and g[5] == 24.0
    # integer wrap-around, exactly as arrow's product
    big = pa.record_batch({"k": pa.array([0] * 70 + [1] * 3, pa.int64()), "v": pa.array([3] * 70 + [-7, 11, 13], pa.int64())})
    gb2, ora2 = pab.GroupBy("k", big), orc.OracleGroupBy(big, "k")
    assert gb2.product("v").equals(ora2.agg("product", "v"))
    assert gb2.product("v").to_pylist()[1] == -1001
    # empty input
    e = pa.record_batch({"k": pa.array([], pa.int64()), "v": pa.array([], pa.float64())})
    ge = pab.GroupBy("k", e)
    r = ge.aggregate(e.column("v"), S2)
    assert all(len(r[a]) == 0 for a in S2) and r["variance"].type == pa.float64()


def test_stage2_resampler(pab, orc):
    rng = np.random.default_rng(3)
    n = 50_000
    ts = np.cumsum(rng.integers(1, 2_000_000_000, n)).astype(np.int64) + 1_577_836_800_000_000_000
    rb = pa.record_batch({"v": pa.array(rng.normal(size=n)), "q": pa.array(rng.integers(1, 4, n), pa.int64())})
    idx = pa.array(ts, pa.timestamp("ns"))
    rs = pab.resample(rb, idx, 60_000_000_000)
    labels = orc.resample_labels(idx, 60_000_000_000)
    ora = orc.OracleGroupBy(pa.record_batch({"k": labels, "v": rb.column("v"), "q": rb.column("q")}), "k")
    assert rs.index().cast(pa.int64()).equals(ora.unique().cast(pa.int64()))
    var = rs.variance()
    want = _numeric(ora, "variance", "v")
    np.testing.assert_allclose(var["v"].to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False), rtol=1e-12, atol=1e-13)
    assert rs.product("q").equals(ora.agg("product", "q"))


def test_stage2_refused_on_merged_handles(pab):
    import torch
    from pandasarrow_b200._lib import PA_PARTIAL_WORDS as W
    rb = pa.record_batch({"k": pa.array([1, 2, 1], pa.int64()), "v": pa.array([1.0, 2.0, 3.0])})
    g = pab.GroupBy("k", rb)
    g.aggregate(rb.column("v"), ["sum", "count"], fetch=False)
    c = g.partials_count(1)
    buf = torch.empty((max(sum(c), 1), W), dtype=torch.int64, device="cuda")
    g.partials_export(1, buf.data_ptr(), buf.shape[0])
    m = pab.MergedGroupBy(buf.data_ptr(), c, ["sum", "count"], "g", "l")
    with pytest.raises(pab.PaError, match="merged"):
        m.aggregate(rb.column("v"), ["variance"])


# ---------------- all / any on boolean columns (GROUPBY_NUMERIC_AGG(all|any, bool), dataframe.cpp:1522-1524) ----------------
@pytest.mark.parametrize("n,G,kw", [(50_000, 7, {}), (200_000, 900, {}), (200_000, 5000, {}), (300_000, 100_000, {"path": "global"})])
def test_bool_all_any_match_oracle(pab, orc, n, G, kw):
    from util import align_to
    rng = np.random.default_rng(n + G)
    k = rng.integers(0, G, n)
    b = rng.random(n) < 0.97                       # mostly true, so that `all` is true for some groups
    b[k % 5 == 0] = True
    b[k % 7 == 1] = False                          # groups that are all false: `any` false
    vm = rng.random(n) < 0.1
    vm[k % 11 == 3] = True                         # all-null groups
    rb = pa.record_batch({"k": pa.array(k, pa.int64(), mask=rng.random(n) < 0.01), "b": pa.array(b, pa.bool_(), mask=vm)})
    gb, ora = pab.GroupBy("k", rb, **kw), orc.OracleGroupBy(rb, "k")
    perm = pa.array(align_to([(x,) for x in gb.unique().to_pylist()], [(x,) for x in ora.unique().to_pylist()]))
    r = gb.aggregate(rb.column("b"), ["all", "any", "count"])
    for a in ("all", "any"):
        want, valid = ora.agg(a, "b", nthreads=8, with_validity=True)
        got = r[a].take(perm)
        assert got.type == pa.bool_()
        assert got.is_valid().to_pylist() == valid.to_pylist(), a            # null for an all-null group ...
        assert got.fill_null(False).to_pylist() == want.to_pylist(), a       # ... which the reference's wrapper turns into false
    assert r["count"].take(perm).equals(ora.agg("count", "b", nthreads=8))
    assert len(set(r["all"].to_pylist())) == 3 and len(set(r["any"].to_pylist())) >= 2     # true, false and null all occur
    # a sliced boolean column (bit offset not a multiple of 8) gives the same answer
    padded = pa.concat_arrays([pa.array([True, None, False], pa.bool_()), rb.column("b")]).slice(3)
    assert gb.aggregate(padded, ["all"])["all"].equals(r["all"])


def test_bool_columns_reject_other_aggregates(pab):
    rb = pa.record_batch({"k": pa.array([1, 2, 1], pa.int64()), "b": pa.array([True, False, True]), "v": pa.array([1.0, 2.0, 3.0])})
    g = pab.GroupBy("k", rb)
    with pytest.raises(pab.PaError, match="boolean columns"):
        g.aggregate(rb.column("b"), ["sum"])
    with pytest.raises(pab.PaError, match="boolean column"):
        g.aggregate(rb.column("v"), ["all"])
    assert g.all("b").to_pylist() == [True, False] and g.any("b").to_pylist() == [True, False]
    with pytest.raises(pab.PaError, match="not part of the last"):
        g.fetch("min")                               # the helper min / max columns are not exposed


# ---------------- count_distinct (GROUPBY_NUMERIC_AGG(count_distinct, int64_t), dataframe.cpp:1528) ----------------
@pytest.mark.parametrize("n,G,kw", [(40_000, 3, {}), (200_000, 700, {}), (200_000, 6000, {}), (300_000, 120_000, {"path": "global"})])
def test_count_distinct_matches_oracle(pab, orc, n, G, kw):
    from util import align_to
    rng = np.random.default_rng(7 * n + G)
    vm = rng.random(n) < 0.1
    k = rng.integers(0, G, n)
    vm[k % 13 == 5] = True                                                  # all-null groups: 0
    f = rng.integers(0, 50, n).astype(np.float64) / 4
    f[rng.random(n) < 0.01] = -0.0
    f[rng.random(n) < 0.01] = np.nan
    rb = pa.record_batch({
        "k": pa.array(k * 17 - 40, pa.int64(), mask=rng.random(n) < 0.01),
        "f": pa.array(f, pa.float64(), mask=vm),                            # few distinct values, -0.0 / 0.0 / NaN among them
        "r": pa.array(rng.normal(size=n), pa.float64(), mask=vm),           # (almost) all distinct
        "i": pa.array(rng.integers(-20, 20, n), pa.int64(), mask=vm),
        "u8": pa.array(rng.integers(0, 9, n).astype(np.uint8), pa.uint8(), mask=vm),
        "f32": pa.array((rng.integers(0, 30, n) / 8).astype(np.float32), pa.float32(), mask=vm),
        "b": pa.array(rng.random(n) < 0.5, pa.bool_(), mask=vm)})
    gb, ora = pab.GroupBy("k", rb, **kw), orc.OracleGroupBy(rb, "k")
    perm = pa.array(align_to([(x,) for x in gb.unique().to_pylist()], [(x,) for x in ora.unique().to_pylist()]))
    for c in ("f", "r", "i", "u8", "f32", "b"):
        got = gb.count_distinct(c).take(perm)
        want = ora.agg("count_distinct", c, nthreads=8)
        assert got.type == pa.int64() and got.null_count == 0
        assert got.equals(want), c
    both = gb.aggregate(rb.column("i"), ["count_distinct", "count", "sum"])
    assert (both["count_distinct"].to_numpy() <= both["count"].to_numpy()).all()


def test_count_distinct_bit_pattern_semantics(pab, orc):
    # arrow's memo table hashes the bits: -0.0 and +0.0 are two values, one NaN payload is one value
    rb = pa.record_batch({"k": pa.array([1, 1, 1, 1, 2, 2, 2, 3], pa.int64()),
                          "v": pa.array([0.0, -0.0, float("nan"), float("nan"), 1.0, 1.0, None, None])})
    g, ora = pab.GroupBy("k", rb), orc.OracleGroupBy(rb, "k")
    assert g.count_distinct("v").to_pylist() == ora.agg("count_distinct", "v").to_pylist() == [3, 1, 0]
    e = pa.record_batch({"k": pa.array([], pa.int64()), "v": pa.array([], pa.float64())})
    assert len(pab.GroupBy("k", e).count_distinct("v")) == 0
