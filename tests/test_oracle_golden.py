"""Pins the oracle against every known-answer vector the reference's own tests hold for the
group-by / resample path.  Citations are /root/reference/tests/<file>:<lines>."""
import datetime as dt

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as orc

IDS = ["allen", "victor", "hannah", "allen", "victor", "hannah", "allen", "victor", "hannah", "allen"]
GENDER = ["male", "female", "male", "male", "female", "male", "male", "female", "male", "male"]
AGE = [16, 10, 10, 20, 30, 40, 15, 25, 35, 45]
HEIGHT = [9, 9, 9, 9, 9, 8, 8, 8, 8, 8]


def people():
    return pa.record_batch({"id": pa.array(IDS), "gender": pa.array(GENDER),
                            "age": pa.array(AGE, pa.int32()), "height": pa.array(HEIGHT, pa.int32())})


def test_single_key_groups_and_order():
    # cudf_examples/dataframe_resample_test.cpp:8-69
    rb = pa.record_batch({"id": pa.array(IDS), "age": pa.array(AGE, pa.int32())})
    g = orc.OracleGroupBy(rb, "id", index=pa.array(range(10), pa.int64()), materialize=True)
    assert g.num_groups == 3
    assert g.unique().to_pylist() == ["allen", "victor", "hannah"]
    assert g.group_slice("age", 0).to_pylist() == [16, 20, 15, 45]
    assert g.group_slice("age", 1).to_pylist() == [10, 30, 25]
    assert g.group_slice("age", 2).to_pylist() == [10, 40, 35]
    assert g.group_slice("id", 0).to_pylist() == ["allen"] * 4
    assert g.row_ids().to_pylist() == [0, 1, 2, 0, 1, 2, 0, 1, 2, 0]


@pytest.mark.parametrize("materialize", [False, True])
def test_gender_aggregates(materialize):
    # cudf_examples/dataframe_resample_test.cpp:71-250
    g = orc.OracleGroupBy(people(), "gender", index=pa.array(range(10), pa.int64()), materialize=materialize)
    assert g.num_groups == 2
    assert g.unique().to_pylist() == ["male", "female"]
    assert g.group_slice("id", 0).to_pylist() == ["allen", "hannah", "allen", "hannah", "allen", "hannah", "allen"]
    assert g.group_slice("age", 0).to_pylist() == [16, 10, 20, 40, 15, 35, 45]
    assert g.group_slice("age", 1).to_pylist() == [10, 30, 25]
    # :166-179 mean, exact doubles
    assert g.agg("mean", "age").to_pylist() == [25.857142857142858, 21.666666666666668]
    assert g.agg("mean", "height").to_pylist() == [8.428571428571429, 8.666666666666666]
    # :181-206 min_max keeps int32
    mn, mx = g.min_max("age")
    assert mn.type == pa.int32() and mn.to_pylist() == [10, 10] and mx.to_pylist() == [45, 30]
    mn, mx = g.min_max("height")
    assert mn.to_pylist() == [8, 8] and mx.to_pylist() == [9, 9]
    # :208-228
    assert g.agg("max", "age").to_pylist() == [45, 30]
    assert g.agg("max", "height").to_pylist() == [9, 9]
    assert g.agg("min", "age").to_pylist() == [10, 10]
    # :230-243 sum of int32 is int64
    s = g.agg("sum", "age")
    assert s.type == pa.int64() and s.to_pylist() == [181, 65]
    assert g.agg("sum", "height").to_pylist() == [59, 26]
    # :245-250
    c = g.agg("count", "age")
    assert c.type == pa.int64() and c.to_pylist() == [7, 3]


def test_bardata_ohlc():
    # cudf_examples/dataframe_resample_test.cpp:252-305 (asserts shapes; values follow from semantics)
    day = pa.array([1, 1, 2, 2, 5], pa.int64())
    rb = pa.record_batch({
        "high": pa.array([11.1, 20.2, 21.0, 15, 20], pa.float32()),
        "low": pa.array([9.1, 9.2, 10.0, 5, 10], pa.float32()),
        "close": pa.array([10.1, 15.2, 20.0, 15, 15], pa.float32()),
        "open": pa.array([10, 20.2, 10.0, 15, 10], pa.float32()),
        "volume": pa.array([100, 200, 210, 1, 2], pa.uint64()),
        "day": day})
    g = orc.OracleGroupBy(rb, "day", materialize=True)
    assert g.num_groups == 3 and g.unique().to_pylist() == [1, 2, 5]
    assert [len(g.group_slice("high", j)) for j in range(3)] == [2, 2, 1]
    f32 = lambda xs: [float(np.float32(x)) for x in xs]
    assert g.agg("first", "open").to_pylist() == f32([10, 10.0, 10])
    assert g.agg("last", "close").to_pylist() == f32([15.2, 15, 15])
    assert g.agg("max", "high").to_pylist() == f32([20.2, 21.0, 20])
    assert g.agg("min", "low").to_pylist() == f32([9.1, 5, 10])
    v = g.agg("sum", "volume")
    assert v.type == pa.uint64() and v.to_pylist() == [300, 211, 2]


def test_apply_sums():
    # dataframe_iterator_test.cpp:11-76
    rb = pa.record_batch({"a": pa.array([1, 1, 3, 1, 1, 1, 3, 8, 2, 2], pa.int32()),
                          "b": pa.array([10, 9, 8, 7, 6, 5, 4, 3, 2, 1], pa.int32())})
    g = orc.OracleGroupBy(rb, "a", materialize=True)
    assert g.num_groups == 4
    assert g.unique().to_pylist() == [1, 3, 8, 2]
    a, b = g.agg("sum", "a").to_pylist(), g.agg("sum", "b").to_pylist()
    assert a == [5, 6, 8, 4] and b == [37, 12, 3, 3]
    assert [x + y for x, y in zip(a, b)] == [42, 18, 11, 7]


def _minute_index(n, start=dt.datetime(2000, 1, 1)):
    base = int((start - dt.datetime(1970, 1, 1)).total_seconds()) * 10**9
    return pa.array([base + i * 60 * 10**9 for i in range(n)], pa.timestamp("ns"))


def _ts(s):
    return dt.datetime.strptime(s, "%Y-%m-%d %H:%M:%S")


THREE_MIN = 3 * 60 * 10**9


@pytest.mark.parametrize("closed_right,label_right,labels,sums", [
    # series_resample_test.cpp:17-30
    (False, False, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00"], [3, 12, 21]),
    # :32-47
    (False, True, ["2000-01-01 00:03:00", "2000-01-01 00:06:00", "2000-01-01 00:09:00"], [3, 12, 21]),
    # :49-69
    (True, True, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00", "2000-01-01 00:09:00"],
     [0, 6, 15, 15]),
])
def test_resample_series(closed_right, label_right, labels, sums):
    idx = _minute_index(9)
    rb = pa.record_batch({"v": pa.array(range(9), pa.int64())})
    g = orc.resample(rb, idx, THREE_MIN, closed_right=closed_right, label_right=label_right)
    assert [t.replace(tzinfo=None) for t in g.unique().to_pylist()] == [_ts(s) for s in labels]
    assert g.agg("sum", "v").to_pylist() == sums


def test_resample_apply_equivalent():
    # series_resample_test.cpp:72-85 — apply(sum + 5) per bucket
    idx = _minute_index(9)
    g = orc.resample(pa.record_batch({"v": pa.array(range(9), pa.int64())}), idx, THREE_MIN)
    assert [s + 5 for s in g.agg("sum", "v").to_pylist()] == [8, 17, 26]


@pytest.mark.parametrize("right,labels,sums", [
    # series_resample_test.cpp:91-107
    (False, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00"], [3, 12, 21]),
    # :109-130
    (True, ["2000-01-01 00:00:00", "2000-01-01 00:03:00", "2000-01-01 00:06:00", "2000-01-01 00:09:00"],
     [0, 6, 15, 15]),
])
def test_downsample_3T(right, labels, sums):
    idx = _minute_index(9)
    lab = orc.downsample_labels(idx, 3, "T", closed_label_right=right)
    g = orc.OracleGroupBy(pa.record_batch({"i": pa.array(range(9), pa.int64())}), "__resampler_idx__", index=lab)
    assert [t.replace(tzinfo=None) for t in g.unique().to_pylist()] == [_ts(s) for s in labels]
    assert g.agg("sum", "i").to_pylist() == sums


def test_downsample_month_mean():
    # series_resample_test.cpp:144-166
    days = [dt.datetime(2018, 1, 7), dt.datetime(2018, 1, 14), dt.datetime(2018, 1, 21), dt.datetime(2018, 1, 28),
            dt.datetime(2018, 2, 4), dt.datetime(2018, 2, 11), dt.datetime(2018, 2, 18), dt.datetime(2018, 2, 25)]
    idx = pa.array(days, pa.timestamp("ns"))
    rb = pa.record_batch({"price": pa.array([10, 11, 9, 13, 14, 18, 17, 19], pa.int32()),
                          "volume": pa.array([50, 60, 40, 100, 50, 100, 40, 50], pa.int32())})
    lab = orc.downsample_labels(idx, 1, "M", closed_label_right=True)
    g = orc.OracleGroupBy(rb, "__resampler_idx__", index=lab)
    assert [t.replace(tzinfo=None) for t in g.unique().to_pylist()] == [dt.datetime(2018, 1, 31), dt.datetime(2018, 2, 28)]
    assert g.agg("mean", "price").to_pylist() == [10.75, 17]
    assert g.agg("mean", "volume").to_pylist() == [62.5, 60]


def test_generate_bins_vectors_via_labels():
    # range_generate_test.cpp:292-310 (commented out upstream): values 0..8 (step 1), edges {0,3,6,9}
    # -> closed-left bins {3,6,9->clipped}.  Restated through resample_labels on ns-resolution ticks.
    idx = pa.array(list(range(9)), pa.timestamp("ns"))
    lab = orc.resample_labels(idx, 3, origin="epoch")
    assert lab.cast(pa.int64()).to_pylist() == [0, 0, 0, 3, 3, 3, 6, 6, 6]
    lab = orc.resample_labels(idx, 3, closed_right=True, label_right=True, origin="epoch")
    assert lab.cast(pa.int64()).to_pylist() == [0, 3, 3, 3, 6, 6, 6, 9, 9]


def test_upsampling_rejected():
    # resample.h:102-105; cudf_examples/dataframe_resample_test.cpp:330-347
    idx = _minute_index(9)
    with pytest.raises(orc.OracleError, match="upSampling"):
        orc.resample_labels(idx, 30 * 10**9)


def test_scalar_aggregations():
    # series_aggregation_test.cpp:45-62,122-149,183-196,277-293
    i32 = lambda v, m=None: pa.array(v, pa.int32(), mask=None if m is None else np.array([not x for x in m]))
    assert orc.scalar_agg(i32([1, 2, 3, 4, 5]), "count").as_py() == 5
    assert orc.scalar_agg(i32([1, 2, 3, 4, 5, 0, 7]), "count").as_py() == 7
    masked = i32([1, 2, 3, 4, 5], [True, True, True, True, False])
    assert orc.scalar_agg(masked, "min").as_py() == 1
    assert orc.scalar_agg(masked, "max").as_py() == 4
    assert orc.scalar_agg(i32([1, 2, 3, 4, 5]), "mean").as_py() == 3.0
    assert orc.scalar_agg(masked, "mean").as_py() == 2.5
    assert orc.scalar_agg(i32([1, 2, 3, 4, 5], [False, True, True, True, True]), "mean").as_py() == 3.5
    assert orc.scalar_agg(i32([1, 2, 3, 4, 5]), "sum").as_py() == 15
    d = pa.array([1, 2, 3, 4, float("nan")], pa.float64(), mask=np.array([False, False, False, False, True]))
    assert orc.scalar_agg(d, "sum", skip_null=True).as_py() == 10
    assert not orc.scalar_agg(d, "sum", skip_null=False).is_valid


def test_semantics_probed_in_survey():
    # SURVEY.md §8c "Oracle semantics established by probe" — pinned here so drift is caught.
    rb = pa.record_batch({
        "k": pa.array([5, None, 5, 7, None, 7, 9], pa.int64()),
        "v": pa.array([1.0, 2.0, None, float("nan"), 4.0, 1.0, None], pa.float64()),
        "i": pa.array([2**62, 1, 2**62, 3, 4, 5, None], pa.int64()),
        "z": pa.array([-0.0, 1.0, 0.0, 1.0, 1.0, 1.0, 1.0], pa.float64())})
    g = orc.OracleGroupBy(rb, "k")
    assert g.unique().to_pylist() == [5, None, 7, 9]          # null key = own group, first-appearance order
    assert g.agg("count", "v").to_pylist() == [1, 2, 2, 0]    # ONLY_VALID
    s = g.agg("sum", "v").to_pylist()
    assert s[0] == 1.0 and s[1] == 6.0 and np.isnan(s[2]) and s[3] is None   # all-null -> null, NaN propagates
    mean, valid = g.agg("mean", "v", with_validity=True)
    assert valid.to_pylist() == [True, True, True, False] and mean.null_count == 0  # validity dropped
    assert g.agg("min", "v").to_pylist()[2] == 1.0 and g.agg("max", "v").to_pylist()[2] == 1.0  # NaN skipped
    assert g.agg("sum", "i").to_pylist()[0] == -2**63          # int64 wraps
    assert g.agg("mean", "i").to_pylist()[0] == float(2**62)   # mean accumulates in double
    mn, mx = g.min_max("z")
    assert str(mn[0].as_py()) == "-0.0" and str(mx[0].as_py()) in ("-0.0", "0.0")
    f = g.agg("first", "v").to_pylist(); l = g.agg("last", "v").to_pylist()
    assert f[0] == 1.0 and f[1] == 2.0 and np.isnan(f[2]) and f[3] is None
    assert l[0] is None and l[1] == 4.0 and l[2] == 1.0 and l[3] is None      # positional, nulls not skipped


def test_multi_key_engine():
    # beyond the reference API (one key) but inside its engine: Grouper::Make({int32, dictionary})
    k1 = pa.array([1, 2, 1, 2, 1, None], pa.int32())
    k2 = pa.array(["x", "x", "y", "x", "x", "x"]).dictionary_encode()
    rb = pa.record_batch({"k1": k1, "k2": k2, "v": pa.array([1., 2., 3., 4., 5., 6.])})
    g = orc.OracleGroupBy(rb, ["k1", "k2"])
    assert g.num_groups == 4
    assert g.unique(0).to_pylist() == [1, 2, 1, None]
    assert g.unique(1).to_pylist() == ["x", "x", "y", "x"]
    assert g.agg("sum", "v").to_pylist() == [6.0, 6.0, 3.0, 6.0]


def test_second_stage_aggregates_pinned_to_arrow_scalar_kernels():
    # GROUPBY_AGG(product), GROUPBY_NUMERIC_AGG(variance|stddev|all|any|count_distinct) (dataframe.cpp:1516-1536):
    # the oracle must equal one arrow::compute call per group slice with default options — checked here through
    # pyarrow's bindings of the same kernels (same libarrow), so the GPU parity tests rest on a pinned oracle.
    import pyarrow.compute as pc
    rng = np.random.default_rng(0)
    n = 5000
    k = rng.integers(0, 9, n)
    vm = rng.random(n) < 0.2
    vm[k == 4] = True                                         # an all-null group
    rb = pa.record_batch({
        "k": pa.array(k, pa.int64()),
        "f": pa.array(rng.normal(2.0, 3.0, n), pa.float64(), mask=vm),
        "i": pa.array(rng.integers(-4, 5, n), pa.int32(), mask=vm),
        "p": pa.array(np.exp(rng.uniform(-0.01, 0.01, n)), pa.float64(), mask=vm),
        "b": pa.array(rng.random(n) < 0.9, pa.bool_(), mask=vm),
        "z": pa.array(rng.choice([0.0, -0.0, 1.5, float("nan")], n), pa.float64(), mask=vm)})
    g = orc.OracleGroupBy(rb, "k")
    keys = g.unique().to_pylist()
    assert keys == list(dict.fromkeys(k.tolist()))            # 9 keys in 5000 rows: first-appearance order
    tbl = pa.table(rb)
    for j, key in enumerate(keys):
        sl = tbl.filter(pc.equal(tbl["k"], key))
        for func, col in (("variance", "f"), ("stddev", "f"), ("variance", "i"), ("stddev", "i")):
            want = getattr(pc, func)(sl[col])
            got, valid = g.agg(func, col, with_validity=True)
            assert valid[j].as_py() == want.is_valid
            if want.is_valid:
                assert got[j].as_py() == want.as_py(), (func, col, key)          # bit for bit: same kernel, same slice
            else:
                assert got[j].as_py() == 0.0                                       # `.value` of a null scalar
        for col in ("p", "i"):
            want = pc.product(sl[col])
            got = g.agg("product", col)[j]
            assert got.is_valid == want.is_valid and (not want.is_valid or got.as_py() == want.as_py()), ("product", col, key)
        for func in ("all", "any"):
            want = getattr(pc, func)(sl["b"])
            got, valid = g.agg(func, "b", with_validity=True)
            assert valid[j].as_py() == want.is_valid and got[j].as_py() == (want.as_py() if want.is_valid else False)
        for col in ("z", "i", "b"):
            assert g.agg("count_distinct", col)[j].as_py() == pc.count_distinct(sl[col]).as_py(), ("count_distinct", col, key)
    # -0.0 / +0.0 / NaN are three distinct values for arrow's memo table
    assert max(g.agg("count_distinct", "z").to_pylist()) == 4
