"""Ingest: host Arrow columns -> device once (pa_column_to_device), pageable memory through the pinned
staging pipeline of h2d_copy; IPC blobs as DataFrame::readBinary reads them (dataframe.cpp:757-791).  The device-resident
columns must aggregate to exactly what the host columns do, and both to the oracle.  Needs a GPU: -m gpu."""
import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

ALL = ["sum", "mean", "count", "min", "max", "first", "last"]


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _same(a: dict, b: dict):
    for k in a:
        assert a[k].equals(b[k]), k


def test_pageable_columns_larger_than_a_staging_chunk(pab, orc):
    # 5 000 003 rows x 8 B = 40 MB per column: two staging chunks, the second one ragged; ordinary Arrow heap memory
    from util import compare_all, with_abs
    rng = np.random.default_rng(3)
    n = 5_000_003
    k = pa.array(rng.integers(0, 777, n), pa.int64())
    v = pa.array(rng.standard_normal(n), pa.float64(), mask=rng.random(n) < 0.01)
    frame = {"k": k, "v": v}
    ora = orc.OracleGroupBy(with_abs(frame), "k")
    with pab.GroupBy("k", frame) as gb:
        compare_all(gb, ora, frame, "v", ALL, "pageable 40 MB columns", key_cols=[k])


def test_to_device_matches_host_path_and_survives_reuse(pab):
    rng = np.random.default_rng(4)
    n = 3_000_001
    k = pa.array(rng.integers(-50, 50, n), pa.int32(), mask=rng.random(n) < 0.02)
    v = pa.array(rng.integers(-10**6, 10**6, n), pa.int64(), mask=rng.random(n) < 0.05)
    # slices with an offset that is not a multiple of 8: the ingest keeps the sub-byte bitmap offset
    ks, vs = k.slice(13, n - 100), v.slice(13, n - 100)
    dk, dv = pab.to_device(ks), pab.to_device(vs)
    with pab.GroupBy("k", {"k": ks, "v": vs}) as h, pab.GroupBy("k", {"k": dk, "v": dv}) as d:
        rh, rd = h.aggregate(vs, ALL), d.aggregate(dv, ALL)
        assert h.unique().equals(d.unique())
        _same(rh, rd)
        # the device-resident column is used in place again and again (no re-upload)
        for aggs in (["sum"], ["min", "max"], ["mean", "count"]):
            _same(h.aggregate(vs, aggs), d.aggregate(dv, aggs))


def test_ipc_blob_zero_copy_columns(pab, orc):
    # DataFrame::readBinary: one record batch in an IPC stream; the columns are views into the blob (pageable)
    from util import compare_all, with_abs
    rng = np.random.default_rng(5)
    n = 2_500_000
    rb = pa.record_batch({"sym": pa.array(rng.integers(0, 300, n), pa.int64()), "px": pa.array(rng.random(n) * 100)})
    sink = pa.BufferOutputStream()
    with pa.ipc.new_stream(sink, rb.schema) as w:
        w.write_batch(rb)
    blob = sink.getvalue()
    back = pa.ipc.open_stream(blob).read_all().to_batches()
    assert len(back) == 1
    b = back[0]
    frame = {"sym": b.column("sym"), "px": b.column("px")}
    ora = orc.OracleGroupBy(with_abs(frame), "sym")
    dk, dv = pab.to_device(b.column("sym")), pab.to_device(b.column("px"))
    with pab.GroupBy("sym", frame) as h, pab.GroupBy("sym", {"sym": dk, "px": dv}) as d:
        rh = compare_all(h, ora, frame, "px", ALL, "ipc blob", key_cols=[frame["sym"]])
        assert h.unique().equals(d.unique())
        _same(rh, d.aggregate(dv, ALL))


@pytest.mark.parametrize("n,chunk,G", [(1_000_003, 100_000, 500), (1_000_003, 333_334, 40_000), (250_000, 1 << 30, 1000), (700_001, 11_000, 3)])
def test_chunked_aggregate_of_host_columns(pab, orc, n, chunk, G):
    """pa_groupby_aggregate_chunked (frames larger than the device): chunks aggregated one after the other and merged like
    the row-range shards of the multi-GPU path — against the ORACLE on the whole columns, group order included."""
    from util import abs_scale, align_to, assert_exact, assert_fp_close, first_appearance_order, with_abs
    rng = np.random.default_rng(n + chunk)
    k = pa.array(rng.integers(-G // 2, G - G // 2, n) * 7919, pa.int64())
    v = pa.array(rng.standard_normal(n), pa.float64(), mask=rng.random(n) < 0.03)
    frame = {"k": k, "v": v}
    ora = orc.OracleGroupBy(with_abs(frame), "k")
    m = pab.aggregate_chunked(k, v, ALL, chunk)
    try:
        ours = [(x,) for x in m.unique().to_pylist()]
        theirs = [(x,) for x in ora.unique().to_pylist()]
        assert ours == first_appearance_order([k]), "not in global first-appearance order"
        perm = pa.array(align_to(ours, theirs))
        for a in ALL:
            got = m.fetch(a).take(perm)
            if a == "mean":
                want, valid = ora.agg("mean", "v", nthreads=8, with_validity=True)
                want = pa.array(want.to_numpy(zero_copy_only=False), pa.float64(), mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool))
                assert_fp_close(got, want, "mean", abs_scale(ora, "v", mean=True))
            elif a == "sum":
                assert_fp_close(got, ora.agg("sum", "v", nthreads=8), "sum", abs_scale(ora, "v"))
            else:
                assert_exact(got, ora.agg(a, "v", nthreads=8), a)
    finally:
        m.close()
