"""utf8 / large_utf8 KEY columns on the device (csrc/strkeys.cuh; C ABI format "u" / "U": validity, offsets, bytes):
hashed, grouped, verified byte for byte and returned as strings — against the oracle, which groups the same strings
through arrow::compute::Grouper like the reference (/root/reference/src/dataframe.cpp:1579-1591; the reference's
golden tests key on strings: tests/cudf_examples/dataframe_resample_test.cpp:8-69)."""
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


def _words(rng, n_distinct, max_len=24):
    out = set()
    while len(out) < n_distinct:
        ln = int(rng.integers(0, max_len + 1))
        out.add("".join(chr(int(c)) for c in rng.integers(0x61, 0x7B, ln)) + ("é" if ln % 5 == 0 and ln else ""))
    return sorted(out)


@pytest.mark.parametrize("n,G,typ,nulls", [(10, 3, pa.string(), False), (200_003, 700, pa.string(), True),
                                           (500_000, 40_000, pa.large_string(), True), (300_000, 1, pa.string(), False)])
def test_utf8_keys_match_oracle(pab, orc, n, G, typ, nulls):
    from util import compare_all, with_abs
    rng = np.random.default_rng(n + G)
    words = _words(rng, G)                       # includes the empty string for G large enough
    idx = rng.integers(0, G, n)
    mask = (rng.random(n) < 0.01) if nulls else None
    k = pa.array([words[i] for i in idx], typ, mask=mask)
    v = pa.array(rng.normal(size=n), mask=rng.random(n) < 0.05)
    w = pa.array(rng.integers(-1000, 1000, n), pa.int64())
    rb = pa.record_batch({"k": k, "v": v, "w": w})
    gb = pab.GroupBy("k", rb)
    ora = orc.OracleGroupBy(with_abs(rb), "k")
    assert gb.unique().type == typ
    assert gb.groupSize() == len(set(k.to_pylist()))
    compare_all(gb, ora, rb, "v", ALL, f"utf8 n={n} G={G} v", key_cols=[rb.column("k")])
    compare_all(gb, ora, rb, "w", ALL, f"utf8 n={n} G={G} w", key_cols=[rb.column("k")])
    gb.close(); ora.close()


def test_utf8_keys_sliced_empty_and_all_null(pab):
    k = pa.array(["x", "bb", "", "bb", None, "x", "", "ccc"]).slice(1, 6)        # bb "" bb None x ""
    v = pa.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    g = pab.GroupBy("k", {"k": k, "v": v})
    assert g.unique().to_pylist() == ["bb", "", None, "x"]
    assert g.sum("v").to_pylist() == [4.0, 8.0, 4.0, 5.0]
    assert g.count("v").to_pylist() == [2, 2, 1, 1]
    e = pab.GroupBy("k", {"k": pa.array([], pa.string()), "v": pa.array([], pa.float64())})
    assert e.groupSize() == 0 and e.unique().to_pylist() == []
    a = pab.GroupBy("k", {"k": pa.array([None, None], pa.string()), "v": pa.array([1.0, 2.0])})
    assert a.unique().to_pylist() == [None] and a.sum("v").to_pylist() == [3.0]


def test_utf8_keys_row_ids_and_groupings(pab):
    rng = np.random.default_rng(5)
    words = _words(rng, 50)
    idx = rng.integers(0, 50, 20_000)
    k = pa.array([words[i] for i in idx])
    g = pab.GroupBy("k", {"k": k})
    uniq = g.unique().to_pylist()
    ids = g.row_ids().to_numpy()
    assert [uniq[i] for i in ids[:2000]] == k.to_pylist()[:2000]
    offsets, rows = g.groupings()
    off, r = offsets.to_numpy(), rows.to_numpy()
    for gi in (0, 7, 49):
        assert all(k[int(x)].as_py() == uniq[gi] for x in r[off[gi]:off[gi + 1]][:50])


def test_utf8_keys_device_resident_buffers(pab):
    """format "u" with CUDA buffers: offsets + bytes already on the device are borrowed, not copied."""
    import ctypes as C
    import torch
    from pandasarrow_b200 import _lib
    from pandasarrow_b200.groupby import _check, _import
    L = _lib.load()
    k = pa.array(["ab", "c", "ab", "ddd", "c", "ab"])
    bufs = k.buffers()
    offs = torch.frombuffer(bufs[1], dtype=torch.int32)[: len(k) + 1].cuda()
    data = torch.frombuffer(bufs[2], dtype=torch.uint8).cuda()
    v = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0], dtype=torch.float64, device="cuda")
    dev, sch = _lib.ArrowDeviceArray(), _lib.ArrowSchema()
    ptrs = (C.c_void_p * 3)(None, offs.data_ptr(), data.data_ptr())
    dev.array.length, dev.array.n_buffers = len(k), 3
    dev.array.buffers = C.cast(ptrs, C.POINTER(C.c_void_p))
    dev.device_id, dev.device_type = 0, _lib.ARROW_DEVICE_CUDA
    sch.format, sch.name, sch.flags = b"u", b"", 2
    opt = _lib.PaOptions()
    L.pa_options_init(C.byref(opt))
    h = C.c_void_p()
    _check(L.pa_groupby_create(C.byref(dev), C.byref(sch), 1, C.byref(opt), C.byref(h)))
    a, s = _lib.ArrowArray(), _lib.ArrowSchema()
    _check(L.pa_groupby_unique(h, 0, C.byref(a), C.byref(s)))
    assert _import(a, s).to_pylist() == ["ab", "c", "ddd"]
    gb = pab.GroupBy.__new__(pab.GroupBy)
    gb._L, gb._h, gb._frame, gb._dicts, gb.key_names, gb._key_args = L, h, {}, [None], ["k"], []
    assert gb.aggregate(pab.DeviceColumn.from_torch(v), ["sum"])["sum"].to_pylist() == [10.0, 7.0, 4.0]
    gb.close()
