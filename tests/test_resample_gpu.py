"""Time-bucket specialisation (pd::resample) on the GPU vs the oracle's restatement of
resample.cpp / resample.h.  Needs a GPU: -m gpu."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

OHLC = ["first", "max", "min", "last", "sum"]
ALL = ["sum", "mean", "count", "min", "max", "first", "last"]
MIN = 60 * 10**9


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _minute_index(n):
    base = int((dt.datetime(2000, 1, 1) - dt.datetime(1970, 1, 1)).total_seconds()) * 10**9
    return pa.array([base + i * MIN for i in range(n)], pa.timestamp("ns"))


def _ts(s):
    return dt.datetime.strptime(s, "%Y-%m-%d %H:%M:%S")


@pytest.mark.parametrize("closed_right,label_right,labels,sums", [
    # /root/reference/tests/series_resample_test.cpp:17-30, :32-47, :49-69
    (False, False, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00"], [3, 12, 21]),
    (False, True, ["2000-01-01 00:03:00", "2000-01-01 00:06:00", "2000-01-01 00:09:00"], [3, 12, 21]),
    (True, True, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00", "2000-01-01 00:09:00"],
     [0, 6, 15, 15]),
])
def test_reference_golden_resample(pab, closed_right, label_right, labels, sums):
    idx = _minute_index(9)
    r = pab.resample({"v": pa.array(range(9), pa.int64())}, idx, 3 * MIN, closed_right=closed_right, label_right=label_right)
    got = r.sum()
    assert [t.replace(tzinfo=None) for t in r.index().to_pylist()] == [_ts(s) for s in labels]
    assert got["v"].type == pa.int64() and got["v"].to_pylist() == sums
    # series_resample_test.cpp:72-85: apply(sum + 5)
    if not closed_right and not label_right:
        assert [s + 5 for s in got["v"].to_pylist()] == [8, 17, 26]


def _compare(pab, orc, ts, frame, freq, aggs, **kw):
    idx = pa.array(ts, pa.timestamp("ns"))
    return _compare_impl(pab, orc, frame, aggs, lambda fr: pab.resample(fr, idx, freq, **kw), lambda rb: orc.resample(rb, idx, freq, **kw))


def _compare_impl(pab, orc, frame, aggs, make_ours, make_oracle):
    """Bucket ORDER: the CUDA path emits buckets in time order (= strict first appearance).  The
    reference inherits arrow::compute::Grouper's id order, which may swap buckets that first appear
    inside the same internal mini-batch, so results are aligned by label before comparing.  When the
    reference's own bin generation throws (resample.cpp:28-41), the CUDA path must fail as well."""
    from util import abs_scale, assert_exact, assert_fp_close, with_abs
    rb = with_abs(frame)
    try:
        ora = make_oracle(rb)
    except orc.OracleError as e:
        with pytest.raises(pab.PaError):
            make_ours(frame).sum()
        return None
    r = make_ours(frame)
    assert r.groupSize() == ora.num_groups
    ours = r.index().cast(pa.int64()).to_numpy()
    theirs = ora.unique().cast(pa.int64()).to_numpy()
    assert r.index().type == ora.unique().type
    assert (np.diff(ours) > 0).all(), "buckets must come out in time order"
    assert np.array_equal(ours, np.sort(theirs)), "bucket labels differ"
    perm = pa.array(np.searchsorted(ours, theirs))
    for name in frame:
        col = frame[name]
        res = r.aggregate(col, aggs)
        assert r.timing()["path"] == "resample"
        for a in aggs:
            got = res[a].take(perm)
            if a == "mean":
                want, valid = ora.agg("mean", name, nthreads=8, with_validity=True)
                want = pa.array(want.to_numpy(zero_copy_only=False), pa.float64(), mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool))
                assert_fp_close(got, want, f"{name} mean", abs_scale(ora, name, mean=True))
            elif a == "sum" and pa.types.is_floating(col.type):
                assert_fp_close(got, ora.agg("sum", name, nthreads=8), f"{name} sum", abs_scale(ora, name))
            else:
                assert_exact(got, ora.agg(a, name, nthreads=8), f"{name} {a}")
    return r


@pytest.mark.parametrize("closed_right", [False, True])
@pytest.mark.parametrize("label_right", [False, True])
@pytest.mark.parametrize("origin", ["start_day", "epoch", "start", "end", "end_day"])
def test_resample_options(pab, orc, closed_right, label_right, origin):
    from pandasarrow_b200 import hostgen as hg
    n = 200_000
    ts = hg.timestamps(n, step_ns=3_000_000_000)   # ~20 ticks per minute
    rng = np.random.default_rng(1)
    frame = {"px": pa.array(rng.random(n) * 100), "qty": pa.array(rng.integers(0, 1000, n), pa.int64())}
    _compare(pab, orc, ts, frame, MIN, OHLC + ["count", "mean"], closed_right=closed_right, label_right=label_right,
             origin=origin, offset_ns=7 * 10**9 if origin == "epoch" else 0)


@pytest.mark.parametrize("n,step,freq", [(1, 1000, MIN), (31, 10**9, MIN), (33, 10**9, 5 * 10**9), (1025, 10**8, 10**9),
                                         (5000, 60_000, MIN), (300_001, 60_000, MIN), (100_000, 10**9, 3600 * 10**9)])
def test_config4_ticks_ohlc(pab, orc, n, step, freq):
    # config 4: sorted timestamp[ns], fixed-width buckets, OHLC + sum on one fp64 column
    from pandasarrow_b200 import hostgen as hg
    ts = hg.timestamps(n, step_ns=step)
    frame = {"px": pa.array(hg.vals(n))}
    _compare(pab, orc, ts, frame, freq, OHLC)


def test_resample_gaps_nulls_and_types(pab, orc):
    rng = np.random.default_rng(9)
    n = 120_000
    # irregular: bursts and long gaps -> many empty buckets (which must not appear), duplicates in the index
    gaps = rng.choice([1, 10**6, 10**9, 400 * 10**9], size=n, p=[0.3, 0.4, 0.29, 0.01])
    ts = 1_600_000_000 * 10**9 + np.cumsum(gaps)
    ts[1000:1100] = ts[1000]
    frame = {"f": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.2),
             "f32": pa.array(rng.standard_normal(n).astype(np.float32)),
             "i": pa.array(rng.integers(-50, 50, n), pa.int32(), mask=rng.random(n) < 0.1),
             "u": pa.array(rng.integers(0, 2**40, n).astype(np.uint64))}
    _compare(pab, orc, ts, frame, MIN, ALL)


def test_resample_errors(pab):
    idx = _minute_index(9)
    with pytest.raises(pab.PaError, match="upSampling"):       # resample.h:102-105
        pab.resample({"v": pa.array(range(9), pa.int64())}, idx, 30 * 10**9)
    with pytest.raises(pab.PaError, match="TimestampArray"):   # resample.cpp:213-216
        pab.resample({"v": pa.array([1.0, 2.0])}, pa.array([1.5, 2.5]), MIN)
    with pytest.raises(pab.PaError, match="positive"):
        pab.resample({"v": pa.array(range(9), pa.int64())}, idx, 0)
    ts = np.arange(5000, dtype=np.int64) * 10**9
    ts[2500], ts[2600] = ts[2600], ts[2500]                    # unsorted in the middle
    r = pab.resample({"v": pa.array(np.ones(5000))}, pa.array(ts, pa.timestamp("ns")), MIN)
    with pytest.raises(pab.PaError, match="sorted"):
        r.sum()
    empty = pab.resample({"v": pa.array([], pa.float64())}, pa.array([], pa.timestamp("ns")), MIN)
    assert empty.groupSize() == 0 and len(empty.sum()["v"]) == 0


# ---------------- DataFrame::downsample: labels computed on the device (csrc/temporal.cuh) ----------------
@pytest.mark.parametrize("rule,closed_label_right,wsm,start_epoch", [
    ("3T", True, True, True), ("3T", False, True, True), ("7T", True, True, False), ("7T", False, True, True),
    ("1H", True, True, True), ("5H", False, True, True), ("1D", True, True, True), ("3D", False, True, True), ("3D", True, True, False),
    ("250L", False, True, True), ("15S", True, True, True),
    ("1W", True, True, True), ("1W", False, False, True), ("2W", True, True, False), ("2W", False, True, True),
    ("1M", True, True, True), ("1M", False, True, True), ("5M", True, True, False), ("2Q", False, True, True), ("1Q", True, True, True),
    ("1Y", True, True, True), ("3Y", False, True, True)])
def test_downsample_labels_on_device_vs_arrow(pab, orc, rule, closed_label_right, wsm, start_epoch):
    """Per-row labels are arrow's Floor/CeilTemporal (minus one day for W / M / Q / Y), as the reference computes them on
    the host; group set, labels and aggregates must match the oracle's group-by on those labels.  Unsorted index."""
    import re
    from util import assert_exact, assert_fp_close
    rng = np.random.default_rng(len(rule) * 7 + closed_label_right)
    n = 150_000
    span = {"L": 4 * 10**9, "S": 3600 * 10**9, "T": 86400 * 10**9, "H": 40 * 86400 * 10**9, "D": 400 * 86400 * 10**9}.get(rule[-1], 9000 * 86400 * 10**9)
    ts = 1_500_000_000 * 10**9 + rng.integers(-span, span, n)
    ts[:50] = (ts[:50] // (86400 * 10**9)) * 86400 * 10**9                 # some rows exactly on day boundaries
    idx = pa.array(ts, pa.timestamp("ns"), mask=rng.random(n) < 0.002)      # a few null timestamps: their own (null-label) group
    frame = {"px": pa.array(rng.random(n) * 100), "qty": pa.array(rng.integers(0, 1000, n), pa.int64(), mask=rng.random(n) < 0.05)}
    mult, unit = re.fullmatch(r"(\d+)([A-Z])", rule).groups()
    labels = orc.downsample_labels(idx, int(mult), unit, closed_label_right, wsm, start_epoch)
    ora = orc.OracleGroupBy(pa.record_batch(frame), "__resampler_idx__", index=labels)
    r = pab.downsample(frame, idx, rule, closed_label_right, wsm, start_epoch)
    assert r.groupSize() == ora.num_groups
    ours, theirs = r.index(), ora.unique()
    assert ours.type == theirs.type
    so = np.argsort(ours.cast(pa.int64()).fill_null(-2**63).to_numpy(), kind="stable")
    st = np.argsort(theirs.cast(pa.int64()).fill_null(-2**63).to_numpy(), kind="stable")
    assert ours.take(pa.array(so)).equals(theirs.take(pa.array(st))), "labels differ"
    for name in frame:
        res = r.aggregate(frame[name], ALL)
        for a in ALL:
            got = res[a].take(pa.array(so))
            want = ora.agg(a, name, nthreads=8).take(pa.array(st))
            if a == "mean":
                m, valid = ora.agg("mean", name, nthreads=8, with_validity=True)
                want = pa.array(m.to_numpy(zero_copy_only=False), pa.float64(), mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool)).take(pa.array(st))
            if a in ("sum", "mean") and pa.types.is_floating(got.type):
                assert_fp_close(got, want, f"{rule} {name} {a}")
            else:
                assert_exact(got, want, f"{rule} {name} {a}")


@pytest.mark.parametrize("unit,per", [("us", 10**3), ("ms", 10**6), ("s", 10**9)])
def test_resample_non_nanosecond_index(pab, orc, unit, per):
    """freq / offset arrive in nanoseconds, the index may tick in s / ms / us: buckets must be the same as for the
    nanosecond twin of the index (the reference itself only handles [ns]; a width that is no whole number of ticks is refused)."""
    from util import assert_exact
    rng = np.random.default_rng(per % 97)
    n = 50_000
    ts_ns = (1_600_000_000 * 10**9 + np.cumsum(rng.integers(1, 40, n)) * 10**9)       # whole seconds: exact in every unit
    frame = {"v": pa.array(rng.integers(0, 100, n), pa.int64())}
    want = pab.resample(frame, pa.array(ts_ns, pa.timestamp("ns")), 5 * MIN, offset_ns=30 * 10**9)
    got = pab.resample(frame, pa.array(ts_ns // per, pa.timestamp(unit)), 5 * MIN, offset_ns=30 * 10**9)
    assert got.groupSize() == want.groupSize()
    assert np.array_equal(got.index().cast(pa.int64()).to_numpy() * per, want.index().cast(pa.int64()).to_numpy())
    assert got.index().type == pa.timestamp(unit)
    assert_exact(got.sum()["v"], want.sum()["v"], unit)
    ora = orc.resample(pa.record_batch(frame), pa.array(ts_ns, pa.timestamp("ns")), 5 * MIN, offset_ns=30 * 10**9)
    assert np.array_equal(np.sort(ora.unique().cast(pa.int64()).to_numpy()), want.index().cast(pa.int64()).to_numpy())
    if per > 1:
        with pytest.raises(pab.PaError, match="whole multiples"):
            pab.resample(frame, pa.array(ts_ns // per, pa.timestamp(unit)), 1500 if unit != "us" else 1500 + 1)


# ---------------- DateOffset (calendar) rules: makeGroupInfo's DateOffset branch, resample.cpp:248-267 ----------------
def _calendar(pab, orc, ts, frame, code, mult, aggs, label_right=False, closed_right=True):
    import pandasarrow_b200.groupby as G
    idx = pa.array(ts, pa.timestamp("ns"))
    rule = f"{mult}{code}"
    return _compare_impl(pab, orc, frame, aggs,
                         lambda fr: G.resample_calendar(fr, idx, rule, closed_right=closed_right, label_right=label_right),
                         lambda rb: orc.resample_calendar(rb, idx, code, mult, closed_right=closed_right, label_right=label_right))


@pytest.mark.parametrize("code,mult", [("D", 1), ("D", 3), ("WS", 1), ("WS", 2), ("MS", 1), ("MS", 5), ("YS", 1)])
@pytest.mark.parametrize("label_right", [False, True])
def test_calendar_rules_vs_oracle(pab, orc, code, mult, label_right):
    # ~3.3 years of irregular ticks (bursts, gaps of days, duplicates, ticks exactly at midnight)
    rng = np.random.default_rng(sum(map(ord, code)) * 31 + mult)
    n = 150_000
    gaps = rng.choice([1, 10**9, 3600 * 10**9, 86400 * 10**9, 9 * 86400 * 10**9], size=n, p=[0.2, 0.5, 0.288, 0.01, 0.002])
    t0 = int(dt.datetime(2019, 11, 28, 17, 5, tzinfo=dt.timezone.utc).timestamp()) * 10**9
    ts = t0 + np.cumsum(gaps)
    day = 86400 * 10**9
    ts[5000:5003] = (ts[5000] // day) * day            # exactly midnight: belongs to the bucket that ENDS there
    ts = np.sort(ts)
    frame = {"px": pa.array(rng.random(n) * 100), "qty": pa.array(rng.integers(0, 1000, n), pa.int64(), mask=rng.random(n) < 0.05)}
    r = _calendar(pab, orc, ts, frame, code, mult, ALL, label_right=label_right)
    assert r is not None and r.groupSize() > 3


def test_calendar_quarter_start(pab, orc):
    # date_range only accepts a quarter rule whose first bin starts in January / February (core.cpp:247-250):
    # first tick in Q2 -> first - 1 quarter = January 1st: accepted; first tick in Q3 -> April 1st: rejected
    day = 86400 * 10**9
    ok0 = int(dt.datetime(2021, 5, 10, tzinfo=dt.timezone.utc).timestamp()) * 10**9
    bad0 = int(dt.datetime(2021, 8, 10, tzinfo=dt.timezone.utc).timestamp()) * 10**9
    n = 20_000
    rng = np.random.default_rng(4)
    frame = {"v": pa.array(rng.standard_normal(n))}
    r = _calendar(pab, orc, ok0 + np.arange(n, dtype=np.int64) * (day // 40), frame, "QS", 1, ALL)
    assert r is not None and r.groupSize() >= 5
    assert _calendar(pab, orc, bad0 + np.arange(n, dtype=np.int64) * (day // 40), frame, "QS", 1, ALL) is None


def test_calendar_errors_like_the_reference(pab, orc):
    import pandasarrow_b200.groupby as G
    day = 86400 * 10**9
    idx = pa.array(1_600_000_000 * 10**9 + np.arange(1000, dtype=np.int64) * day, pa.timestamp("ns"))
    frame = {"v": pa.array(np.arange(1000, dtype=np.float64))}
    for rule, msg in (("M", "MonthEnd not supported"), ("Q", "QuarterEnd not supported"), ("W", "WeekEnd not supported"), ("Y", "YearEnd not supported")):
        with pytest.raises(pab.PaError, match=msg):
            G.resample_calendar(frame, idx, rule, closed_right=True)
        with pytest.raises(orc.OracleError, match=msg):
            orc.resample_labels_calendar(idx, rule, 1, True)
    with pytest.raises(pab.PaError, match="closed_left is not currently supported"):
        G.resample_calendar(frame, idx, "MS", closed_right=False)
    # fewer rows than buckets: upSampling (resample.h:102-105)
    sparse = pa.array(1_600_000_000 * 10**9 + np.arange(5, dtype=np.int64) * 40 * day, pa.timestamp("ns"))
    with pytest.raises(pab.PaError, match="upSampling"):
        G.resample_calendar({"v": pa.array(np.arange(5.0))}, sparse, "D", closed_right=True)
    with pytest.raises(orc.OracleError, match="upSampling"):
        orc.resample_labels_calendar(sparse, "D", 1, True)
    us = pa.array(np.arange(1000, dtype=np.int64) * 86400 * 10**6, pa.timestamp("us"))
    with pytest.raises(pab.PaError, match=r"timestamp\[ns\]"):
        G.resample_calendar(frame, us, "MS", closed_right=True)


@pytest.mark.parametrize("nullable", [False, True])
def test_resample_min_max_special_values(pab, orc, nullable):
    """min / max of fp64 values are folded as doubles per lane (ordered compares: a NaN never wins) and order-mapped when a
    run closes: NaN rows, buckets that hold nothing but NaN, +-inf only buckets and signed zeros against the oracle, on the
    steady-state path (no bitmap) and the per-batch path (bitmap)."""
    from pandasarrow_b200 import hostgen as hg
    n = 400_000
    ts = hg.timestamps(n, step_ns=500_000_000)     # ~120 ticks per minute: runs longer than one 128-row iteration
    rng = np.random.default_rng(11)
    v = rng.standard_normal(n) * 50
    v[rng.random(n) < 0.05] = np.nan
    v[rng.random(n) < 0.01] = np.inf
    v[rng.random(n) < 0.01] = -np.inf
    v[rng.random(n) < 0.03] = 0.0
    v[rng.random(n) < 0.03] = -0.0
    b = (np.asarray(ts) - np.asarray(ts)[0]) // MIN
    v[b == 3] = np.nan                 # nothing but NaN
    v[b == 5] = np.inf                 # +inf only
    v[b == 7] = -np.inf                # -inf only
    v[(b == 9) & (np.arange(n) % 2 == 0)] = np.nan
    v[(b == 9) & (np.arange(n) % 2 == 1)] = -np.inf
    mask = (rng.random(n) < 0.1) if nullable else None
    if nullable:
        mask[b == 11] = True           # an all-null bucket
    frame = {"px": pa.array(v, pa.float64(), mask=mask)}
    _compare(pab, orc, ts, frame, MIN, ["min", "max", "sum", "count", "first", "last"])
