"""The oracle's DateOffset branch (oracle_groupby.cpp resample_labels_calendar: std::chrono calendar) against an
independent pure-Python model of the same reference code (datetime.date arithmetic):
/root/reference/src/resample.cpp:248-267 (makeGroupInfo), :180-200 (adjustBinEdges), :11-83 (generate_bins_dt64),
core.cpp:12-60 (DateOffset::add), core.cpp:175-265 (date_range / switchFunction), resample.h:19-43 (downsample)."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

DAY = 86400 * 10**9
EPOCH = dt.date(1970, 1, 1)


def _ns(d: dt.date) -> int:
    return (d - EPOCH).days * DAY


def _add_months(d: dt.date, k: int) -> dt.date:          # only year / month survive in the callers below
    ym = d.year * 12 + (d.month - 1) + k
    return dt.date(ym // 12, ym % 12 + 1, 1)


def _offset_add(d: dt.date, code: str, k: int) -> dt.date:   # core.cpp:12-60
    if code == "D":
        return d + dt.timedelta(days=k)
    if code == "WS":
        return d + dt.timedelta(weeks=k)
    if code == "MS":
        return _add_months(d, k)
    if code == "QS":
        m = _add_months(d, 3 * k)
        return dt.date(m.year, (m.month - 1) // 3 * 3 + 1, 1)
    if code == "YS":
        return dt.date(d.year + k, 1, 1)
    raise ValueError(code)


def model_labels(ts: np.ndarray, code: str, k: int, label_right: bool) -> np.ndarray:
    first = EPOCH + dt.timedelta(days=int(ts.min() // DAY))
    last = EPOCH + dt.timedelta(days=int(ts.max() // DAY))
    start, stop = _offset_add(first, code, -k), _offset_add(last, code, k)
    assert start < stop
    if code == "QS" and start.month // 3 != 0:
        raise RuntimeError("A quarter freq requires month is on a quarter")
    binner, it, i = [], start, 0
    while it <= stop:                                     # day / week / month / year iterator stepping k units
        binner.append(_ns(it))
        i += 1
        it = _offset_add(start, code, i * k) if code in ("MS", "QS", "YS") else (start + dt.timedelta(days=(7 if code == "WS" else 1) * i * k))
    edges = list(binner)
    if not (code == "D" and k == 1):                      # adjustBinEdges
        edges = [e + DAY - 1 for e in edges]
        if edges[-2] > int(ts.max()):
            edges.pop(); binner.pop()
    assert ts[0] >= edges[0] and ts[-1] <= edges[-1]
    bins, j = [], 0                                       # generate_bins_dt64, right closed
    for r in edges[1:]:
        while j < len(ts) and ts[j] <= r:
            j += 1
        bins.append(j)
    labels = binner[1:] if label_right else binner
    labels = labels[:len(bins)]
    if bins[-1] < len(labels):
        raise RuntimeError("upSampling is not implemented.")
    out, prev = np.empty(bins[-1], dtype=np.int64), 0
    for b, lab in zip(bins, labels):
        out[prev:b] = lab
        prev = b
    return out


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _ticks(seed, n=40_000, start=dt.datetime(2019, 12, 30, 22, 15)):
    rng = np.random.default_rng(seed)
    gaps = rng.choice([1, 10**9, 3600 * 10**9, DAY, 11 * DAY], size=n, p=[0.2, 0.4, 0.37, 0.025, 0.005])
    t0 = int((start - dt.datetime(1970, 1, 1)).total_seconds()) * 10**9
    ts = t0 + np.cumsum(gaps)
    ts[100:103] = (ts[100] // DAY) * DAY                  # exactly midnight
    return np.sort(ts)


@pytest.mark.parametrize("code,k", [("D", 1), ("D", 2), ("D", 7), ("WS", 1), ("WS", 3), ("MS", 1), ("MS", 4), ("YS", 1), ("YS", 2)])
@pytest.mark.parametrize("label_right", [False, True])
def test_oracle_calendar_labels_match_the_python_model(orc, code, k, label_right):
    ts = _ticks(sum(map(ord, code)) + k)
    got = orc.resample_labels_calendar(pa.array(ts, pa.timestamp("ns")), code, k, True, label_right)
    want = model_labels(ts, code, k, label_right)
    assert got.type == pa.timestamp("ns")
    assert np.array_equal(got.cast(pa.int64()).to_numpy(), want)


def test_oracle_calendar_quarter_and_leap_february(orc):
    # QS is only accepted when (first - one quarter) starts in January (core.cpp:247-250)
    ok = _ticks(5, start=dt.datetime(2020, 4, 20))       # Q2 2020 - 1 quarter = 2020-01-01
    got = orc.resample_labels_calendar(pa.array(ok, pa.timestamp("ns")), "QS", 1, True, False)
    assert np.array_equal(got.cast(pa.int64()).to_numpy(), model_labels(ok, "QS", 1, False))
    bad = _ticks(6, start=dt.datetime(2020, 7, 20))
    with pytest.raises(orc.OracleError, match="quarter"):
        orc.resample_labels_calendar(pa.array(bad, pa.timestamp("ns")), "QS", 1, True, False)
    # a leap day inside the range: month starts after 2020-02-29
    feb = _ticks(7, n=5000, start=dt.datetime(2020, 2, 27))
    got = orc.resample_labels_calendar(pa.array(feb, pa.timestamp("ns")), "MS", 1, True, True)
    assert np.array_equal(got.cast(pa.int64()).to_numpy(), model_labels(feb, "MS", 1, True))


def test_oracle_calendar_known_answer(orc):
    # hand-computed: hourly ticks 2020-01-31 22:00 .. 2020-02-02 03:00, rule "MS".
    # binner = [2019-12-01, 2020-01-01, 2020-02-01, 2020-03-01]; edges = binner + 1 day - 1 ns, and since
    # edges[-2] = 2020-02-01 23:59:59.999999999 < max the last edge stays.  Bin 0 (label 2019-12-01) = ticks up to
    # 2020-01-01 23:59 -> none; bin 1 (label 2020-01-01) = ticks up to 2020-02-01 23:59:59.999999999 -> 26 ticks;
    # bin 2 (label 2020-02-01) = the remaining 4.  Empty bin 0 yields no rows.
    t0 = int((dt.datetime(2020, 1, 31, 22) - dt.datetime(1970, 1, 1)).total_seconds()) * 10**9
    ts = t0 + np.arange(30, dtype=np.int64) * 3600 * 10**9
    got = orc.resample_labels_calendar(pa.array(ts, pa.timestamp("ns")), "MS", 1, True, False).cast(pa.int64()).to_numpy()
    jan1, feb1 = _ns(dt.date(2020, 1, 1)), _ns(dt.date(2020, 2, 1))
    assert got.tolist() == [jan1] * 26 + [feb1] * 4
