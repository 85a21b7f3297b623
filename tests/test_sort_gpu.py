"""Device argsort / sorted take (pa_sort_create, sort.cuh) against the oracle's restatement of Series::argsort /
Series::sort / DataFrame::sort_index (arrow's array_sort_indices + Take; series.cpp:864-868,978-992,
dataframe.cpp:1062-1071).  Index work: bit-exact.  Needs a GPU: -m gpu."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _check(pab, orc, col, others=()):
    from util import assert_exact                 # (NaN-aware: Array.equals treats NaN != NaN)
    for asc in (True, False):
        with pab.Sorted(col, ascending=asc) as s:
            got = s.indices()
            want = orc.array_sort(col, asc)
            assert got.type == pa.uint64()
            assert got.equals(want), f"indices differ (ascending={asc}, type={col.type})"
            assert_exact(s.take(col), orc.array_sort(col, asc, take=True), f"sorted values (ascending={asc})")
            for o in others:                      # DataFrame::sort_index: the same indices applied to every column
                assert s.take(o).equals(o.take(want))


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 4095, 4096, 4097, 100_003, 1_000_000])
def test_sort_float64_with_ties_nan_and_nulls(pab, orc, n):
    rng = np.random.default_rng(n)
    v = np.round(rng.standard_normal(n) * 50, 1)           # many ties: stability matters
    if n > 10:
        v[rng.integers(0, n, n // 20)] = np.nan
        v[rng.integers(0, n, n // 50)] = -0.0
        v[rng.integers(0, n, n // 50)] = 0.0
        v[rng.integers(0, n, 3)] = np.inf
        v[rng.integers(0, n, 3)] = -np.inf
    mask = rng.random(n) < 0.07
    col = pa.array(v, pa.float64(), mask=mask)
    other = pa.array(rng.integers(-1000, 1000, n), pa.int32(), mask=rng.random(n) < 0.1)
    flags = pa.array(rng.random(n) < 0.5, pa.bool_())
    _check(pab, orc, col, (other, flags))


@pytest.mark.parametrize("typ", [pa.int64(), pa.uint64(), pa.int32(), pa.uint32(), pa.int16(), pa.uint8(), pa.float32(), pa.timestamp("ns")])
def test_sort_value_types(pab, orc, typ):
    rng = np.random.default_rng(5)
    n = 200_001
    if pa.types.is_floating(typ):
        raw = rng.standard_normal(n).astype(np.float32)
    elif pa.types.is_timestamp(typ):
        raw = rng.integers(-10**18, 10**18, n)
    else:
        info = np.iinfo(typ.to_pandas_dtype())
        raw = rng.integers(info.min, info.max, n, dtype=typ.to_pandas_dtype(), endpoint=True)
        raw[:100] = info.min
        raw[100:200] = info.max
    col = pa.array(raw, typ, mask=rng.random(n) < 0.03)
    _check(pab, orc, col)


def test_sort_small_range_skips_constant_digits(pab, orc):
    # values in [0, 1000): six of the seven 10-bit digits are constant and their passes are skipped
    rng = np.random.default_rng(8)
    n = 3_000_000
    col = pa.array(rng.integers(0, 1000, n), pa.int64())
    _check(pab, orc, col, (pa.array(rng.random(n)),))
    const = pa.array(np.full(10_000, 42, dtype=np.int64))     # nothing to sort at all
    _check(pab, orc, const)


def test_sort_sorted_index_like_config4(pab, orc):
    # DataFrame::sort_index on a shuffled timestamp index restores time order; then resample works on it
    from pandasarrow_b200 import hostgen as hg
    n = 500_000
    ts = hg.timestamps(n, step_ns=10**9)
    perm = np.random.default_rng(1).permutation(n)
    shuffled = pa.array(np.asarray(ts)[perm], pa.timestamp("ns"))
    vals = pa.array(hg.vals(n)[perm])
    with pab.Sorted(shuffled) as s:
        idx_sorted, v_sorted = s.take(shuffled), s.take(vals)
    assert idx_sorted.equals(pa.array(ts, pa.timestamp("ns")))
    assert v_sorted.equals(pa.array(hg.vals(n)))


def test_sort_errors(pab):
    with pytest.raises(pab.PaError):
        pab.Sorted(pa.array(["b", "a"]))
