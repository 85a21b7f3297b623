"""Parity of the CUDA path (through the C ABI) against the oracle.  Needs a GPU: -m gpu."""
import os

import numpy as np
import pyarrow as pa
import pytest

pytestmark = pytest.mark.gpu

ALL = ["sum", "mean", "count", "min", "max", "first", "last"]
NARROW = ["sum", "mean", "count", "first"]


@pytest.fixture(scope="module")
def pab():
    import pandasarrow_b200 as p
    return p


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _both(pab, orc, frame, key, **kw):
    from util import with_abs
    rb = frame if isinstance(frame, pa.RecordBatch) else pa.record_batch(frame)
    return pab.GroupBy(key, rb, **kw), orc.OracleGroupBy(with_abs(rb), key), rb


def _cmp(gb, ora, rb, column, aggs, what, keys=("k",)):
    from util import compare_all
    return compare_all(gb, ora, rb, column, aggs, what, key_cols=[rb.column(k) for k in keys])


# ---------------- the reference's own golden vectors, through the C ABI ----------------
def test_reference_golden_people(pab):
    # /root/reference/tests/cudf_examples/dataframe_resample_test.cpp:71-250
    ids = ["allen", "victor", "hannah", "allen", "victor", "hannah", "allen", "victor", "hannah", "allen"]
    gender = ["male", "female", "male", "male", "female", "male", "male", "female", "male", "male"]
    rb = pa.record_batch({"id": pa.array(ids), "gender": pa.array(gender),
                          "age": pa.array([16, 10, 10, 20, 30, 40, 15, 25, 35, 45], pa.int32()),
                          "height": pa.array([9, 9, 9, 9, 9, 8, 8, 8, 8, 8], pa.int32())})
    g = pab.GroupBy("gender", rb)
    assert g.groupSize() == 2
    assert g.unique().to_pylist() == ["male", "female"]
    m = g.mean(["age", "height"])
    assert m["age"].to_pylist() == [25.857142857142858, 21.666666666666668]
    assert m["height"].to_pylist() == [8.428571428571429, 8.666666666666666]
    mm = g.min_max("age")
    assert mm["min"].type == pa.int32() and mm["min"].to_pylist() == [10, 10] and mm["max"].to_pylist() == [45, 30]
    mm = g.min_max(["age", "height"])
    assert mm["height_min"].to_pylist() == [8, 8] and mm["height_max"].to_pylist() == [9, 9]
    assert g.max("age").to_pylist() == [45, 30] and g.min("age").to_pylist() == [10, 10]
    s = g.sum(["age", "height"])
    assert s["age"].type == pa.int64() and s["age"].to_pylist() == [181, 65] and s["height"].to_pylist() == [59, 26]
    c = g.count("age")
    assert c.type == pa.int64() and c.to_pylist() == [7, 3]
    g2 = pab.GroupBy("id", rb)
    assert g2.unique().to_pylist() == ["allen", "victor", "hannah"]   # :38-41 first-appearance order


def test_reference_golden_apply_sums_and_ohlc(pab):
    # dataframe_iterator_test.cpp:11-76
    rb = pa.record_batch({"a": pa.array([1, 1, 3, 1, 1, 1, 3, 8, 2, 2], pa.int32()),
                          "b": pa.array([10, 9, 8, 7, 6, 5, 4, 3, 2, 1], pa.int32())})
    g = pab.GroupBy("a", rb)
    assert g.groupSize() == 4 and g.unique().to_pylist() == [1, 3, 8, 2]
    assert g.sum("a").to_pylist() == [5, 6, 8, 4] and g.sum("b").to_pylist() == [37, 12, 3, 3]
    # cudf_examples/dataframe_resample_test.cpp:252-305
    f32 = lambda xs: [float(np.float32(x)) for x in xs]
    rb = pa.record_batch({
        "high": pa.array([11.1, 20.2, 21.0, 15, 20], pa.float32()), "low": pa.array([9.1, 9.2, 10.0, 5, 10], pa.float32()),
        "close": pa.array([10.1, 15.2, 20.0, 15, 15], pa.float32()), "open": pa.array([10, 20.2, 10.0, 15, 10], pa.float32()),
        "volume": pa.array([100, 200, 210, 1, 2], pa.uint64()), "day": pa.array([1, 1, 2, 2, 5], pa.int64())})
    g = pab.GroupBy("day", rb)
    assert g.unique().to_pylist() == [1, 2, 5]
    assert g.first("open").to_pylist() == f32([10, 10.0, 10]) and g.last("close").to_pylist() == f32([15.2, 15, 15])
    assert g.max("high").to_pylist() == f32([20.2, 21.0, 20]) and g.min("low").to_pylist() == f32([9.1, 5, 10])
    v = g.sum("volume")
    assert v.type == pa.uint64() and v.to_pylist() == [300, 211, 2]


# ---------------- seeded parity against the oracle ----------------
@pytest.mark.parametrize("path", ["auto", "global"])
@pytest.mark.parametrize("n,G", [(1_000_000, 1000), (300_000, 16), (300_000, 256), (500_007, 4096),
                                 (600_000, 65536), (400_000, 250_000)])
def test_config1_2_int64_key_f64_val(pab, orc, path, n, G):
    from pandasarrow_b200 import hostgen as hg
    frame = {"k": pa.array(hg.keys(n, G)), "v": pa.array(hg.vals(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path, expected_groups=G if path == "global" else 0)
    assert gb.groupSize() == ora.num_groups
    _cmp(gb, ora, rb, "v", ["sum", "mean", "count"], f"n={n} G={G} {path} narrow")
    t = gb.timing()
    if path == "auto" and G <= 1000:
        assert t["path"] == "lowcard"
    _cmp(gb, ora, rb, "v", ALL, f"n={n} G={G} {path} all")
    # count invariants: sum of counts = n
    assert sum(gb.count("v").to_pylist()) == n


@pytest.mark.parametrize("path", ["auto", "global"])
@pytest.mark.parametrize("vtype", [pa.float64(), pa.float32(), pa.int64(), pa.int32(), pa.uint64(), pa.uint32(),
                                   pa.int16(), pa.uint8()])
def test_value_types_with_nulls(pab, orc, path, vtype):
    rng = np.random.default_rng(5)
    n, G = 200_003, 300
    k = rng.integers(0, G, n)
    if pa.types.is_floating(vtype):
        v = rng.standard_normal(n) * 1e3
    elif pa.types.is_unsigned_integer(vtype):
        v = rng.integers(0, 200, n)
    else:
        v = rng.integers(-100, 100, n)
    mask = rng.random(n) < 0.1
    # one group is all-null
    mask |= (k == 7)
    frame = {"k": pa.array(k, pa.int64()), "v": pa.array(v, vtype, mask=mask)}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    _cmp(gb, ora, rb, "v", ALL, f"{vtype} {path}")


@pytest.mark.parametrize("path", ["auto", "global"])
def test_null_keys_int32_key_and_sentinel_key(pab, orc, path):
    rng = np.random.default_rng(11)
    n = 100_001
    k = rng.integers(-50, 50, n).astype(np.int32)
    kmask = rng.random(n) < 0.01
    frame = {"k": pa.array(k, pa.int32(), mask=kmask), "v": pa.array(rng.random(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    _cmp(gb, ora, rb, "v", ALL, f"int32 nullable key {path}")
    # a real key equal to the table's empty sentinel, INT64_MIN/MAX, -1, 0
    sentinel = np.array([0x9E3779B97F4A7C15], dtype=np.uint64).view(np.int64)[0]
    special = np.array([sentinel, np.iinfo(np.int64).min, np.iinfo(np.int64).max, -1, 0], dtype=np.int64)
    k = special[rng.integers(0, len(special), n)]
    frame = {"k": pa.array(k), "v": pa.array(rng.random(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    _cmp(gb, ora, rb, "v", ALL, f"sentinel keys {path}")
    # uint64 keys and timestamp keys
    frame = {"k": pa.array(k.view(np.uint64)), "v": pa.array(rng.random(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    _cmp(gb, ora, rb, "v", NARROW, f"uint64 keys {path}")
    frame = {"k": pa.array(np.abs(k) % 1000, pa.timestamp("ns")), "v": pa.array(rng.random(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    _cmp(gb, ora, rb, "v", NARROW, f"timestamp keys {path}")


@pytest.mark.parametrize("path", ["auto", "global"])
def test_nan_inf_and_all_nan_groups(pab, orc, path):
    k = pa.array([1, 1, 2, 2, 3, 3, 4, 4, 5], pa.int64())
    v = pa.array([1.0, float("nan"), float("nan"), float("nan"), float("inf"), 1.0, float("-inf"), float("inf"), None])
    gb, ora, rb = _both(pab, orc, {"k": k, "v": v}, "k", path=path)
    _cmp(gb, ora, rb, "v", ALL, f"nan/inf {path}")


@pytest.mark.parametrize("path", ["auto", "global"])
@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 255, 256, 257, 1023, 4097])
def test_edge_sizes(pab, orc, path, n):
    rng = np.random.default_rng(n)
    frame = {"k": pa.array(rng.integers(0, 5, n), pa.int64()), "v": pa.array(rng.random(n))}
    gb, ora, rb = _both(pab, orc, frame, "k", path=path)
    assert gb.groupSize() == ora.num_groups
    if n:
        _cmp(gb, ora, rb, "v", ALL, f"n={n} {path}")


@pytest.mark.parametrize("path", ["auto", "global"])
def test_sliced_unaligned_inputs(pab, orc, path):
    # Arrow `offset` must be honoured: values and validity bit offsets
    rng = np.random.default_rng(3)
    n = 50_000
    k = pa.array(rng.integers(0, 100, n), pa.int64())
    v = pa.array(rng.random(n), mask=rng.random(n) < 0.2)
    for off in (1, 3, 8, 13):
        frame = {"k": k.slice(off, n - off - 5), "v": v.slice(off, n - off - 5)}
        gb, ora, rb = _both(pab, orc, frame, "k", path=path)
        _cmp(gb, ora, rb, "v", ALL, f"offset {off} {path}")


@pytest.mark.parametrize("path", ["auto", "global"])
def test_config3_multi_key_dictionary_nullable(pab, orc, path):
    # config 3: int32 key x dictionary-encoded string key, nullable fp64 and int64 values, 1 % null keys
    rng = np.random.default_rng(17)
    n = 300_000
    k1 = pa.array(rng.integers(0, 50, n).astype(np.int32), mask=rng.random(n) < 0.01)
    words = np.array([f"sym{i:02d}" for i in range(16)])
    k2 = pa.array(words[rng.integers(0, 16, n)], mask=rng.random(n) < 0.01).dictionary_encode()
    frame = {"k1": k1, "k2": k2, "f": pa.array(rng.random(n), mask=rng.random(n) < 0.1),
             "i": pa.array(rng.integers(-1000, 1000, n), pa.int64(), mask=rng.random(n) < 0.1)}
    rb = pa.record_batch(frame)
    gb = pab.GroupBy(["k1", "k2"], rb, path=path)
    from util import with_abs
    ora = orc.OracleGroupBy(with_abs(rb), ["k1", "k2"])
    assert gb.groupSize() == ora.num_groups
    _cmp(gb, ora, rb, "f", ALL, f"multi-key f64 {path}", keys=("k1", "k2"))
    _cmp(gb, ora, rb, "i", ALL, f"multi-key i64 {path}", keys=("k1", "k2"))


def test_device_resident_inputs_and_generator(pab, orc):
    # zero-copy CUDA buffers (ARROW_DEVICE_CUDA) + device generator == host generator
    import torch
    from pandasarrow_b200 import hostgen as hg
    n, G = 2_000_003, 1000
    k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
    bits = torch.empty((n + 7) // 8, dtype=torch.uint8, device="cuda")
    pab.synth.keys(k, G); pab.synth.vals(v); pab.synth.validity(bits, n)
    torch.cuda.synchronize()
    assert np.array_equal(k.cpu().numpy(), hg.keys(n, G))
    assert np.array_equal(v.cpu().numpy(), hg.vals(n))
    assert np.array_equal(np.unpackbits(bits.cpu().numpy(), bitorder="little")[:n].astype(bool), hg.valid_mask(n))
    nulls = int(n - hg.valid_mask(n).sum())
    frame_h = pa.record_batch({"k": pa.array(hg.keys(n, G)), "v": pa.array(hg.vals(n), mask=~hg.valid_mask(n))})
    ora = orc.OracleGroupBy(frame_h, "k")
    dk = pab.DeviceColumn.from_torch(k)
    dv = pab.DeviceColumn.from_torch(v, valid=bits, null_count=nulls)
    gb = pab.GroupBy("k", {"k": dk, "v": dv})
    res = gb.aggregate(dv, ALL)
    host = pab.GroupBy("k", frame_h).aggregate(frame_h.column("v"), ALL)
    from util import assert_exact, assert_fp_close
    for a in ALL:   # device-resident and host-staged inputs agree (fp sums on the global path: atomics order)
        (assert_fp_close if a in ("sum", "mean") else assert_exact)(res[a], host[a], a)
    narrow = gb.aggregate(dv, NARROW); narrow_h = pab.GroupBy("k", frame_h).aggregate(frame_h.column("v"), NARROW)
    assert gb.timing()["path"] == "lowcard"
    for a in NARROW:   # the shared-memory path is deterministic: bit for bit
        assert narrow[a].equals(narrow_h[a]), a
    _cmp(pab.GroupBy("k", frame_h), ora, frame_h, "v", ALL, "device generator")


def test_run_to_run_determinism_lowcard(pab):
    from pandasarrow_b200 import hostgen as hg
    n, G = 3_000_000, 1000
    rb = pa.record_batch({"k": pa.array(hg.keys(n, G)), "v": pa.array(hg.vals(n))})
    first = None
    for _ in range(3):
        g = pab.GroupBy("k", rb)
        r = g.aggregate(rb.column("v"), ["sum", "mean"])
        assert g.timing()["path"] == "lowcard"
        if first is None:
            first = r
        else:
            assert r["sum"].equals(first["sum"]) and r["mean"].equals(first["mean"])


def test_errors(pab):
    with pytest.raises(RuntimeError):
        pab.GroupBy("nope", pa.record_batch({"k": pa.array([1, 2])}))
    g = pab.GroupBy("k", pa.record_batch({"k": pa.array([1, 2]), "v": pa.array([1.0, 2.0])}))
    with pytest.raises(pab.PaError):
        g.aggregate(pa.array([1.0, 2.0, 3.0]), ["sum"])          # length mismatch
    with pytest.raises(pab.PaError):
        g.fetch("max")                                            # not computed
    with pytest.raises(pab.PaError):
        pab.GroupBy("k", pa.record_batch({"k": pa.array([[1], [2]])}))   # unsupported key type


# ---------------- the shared-memory path's modes: dense addressing, replication, hash, fallbacks ----------------
def _rand_vals(rng, n):
    return pa.array(rng.random(n) * 100.0 - 20.0)


@pytest.mark.parametrize("no_dense", [False, True])
@pytest.mark.parametrize("G,lo", [(16, 0), (16, -7), (200, 10**12), (1000, -500), (1000, 2**40)])
def test_lowcard_dense_windows_and_hash_mode(pab, orc, no_dense, G, lo):
    rng = np.random.default_rng(G + (1 if no_dense else 0))
    n = 400_003
    frame = {"k": pa.array(rng.integers(lo, lo + G, n), pa.int64()), "v": _rand_vals(rng, n)}
    gb, ora, rb = _both(pab, orc, frame, "k", no_dense=no_dense)
    _cmp(gb, ora, rb, "v", ["sum", "mean", "count", "first"], f"G={G} lo={lo} no_dense={no_dense} narrow")
    t = gb.timing()
    assert t["path"] == "lowcard" and t["mode"] == ("hash" if no_dense else "dense") and t["passes"] == 1
    if not no_dense and G == 16:
        assert t["replication"] == 32          # every lane owns its accumulator slots
    _cmp(gb, ora, rb, "v", ALL, f"G={G} lo={lo} no_dense={no_dense} all")
    assert gb.timing()["path"] == "lowcard"


def test_lowcard_scattered_64bit_keys_use_hash_mode(pab, orc):
    rng = np.random.default_rng(23)
    n, G = 500_000, 1000
    pool = rng.integers(np.iinfo(np.int64).min, np.iinfo(np.int64).max, G, dtype=np.int64)
    frame = {"k": pa.array(pool[rng.integers(0, G, n)]), "v": _rand_vals(rng, n)}
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", ["sum", "mean", "count", "first"], "scattered keys")
    t = gb.timing()
    assert t["path"] == "lowcard" and t["mode"] == "hash" and t["passes"] == 1
    a = gb.aggregate(rb.column("v"), ["sum"])["sum"]
    b = pab.GroupBy("k", rb).aggregate(rb.column("v"), ["sum"])["sum"]
    assert a.equals(b)                          # hash mode is run-to-run deterministic too


def _keys_with_home(home, count, seed):
    """int64 keys whose hash-mode home slot (top 12 bits of lc_mix(key), lowcard.cuh) is `home`."""
    C = 0x9E3779B97F4A7C15
    cinv = pow(C, -1, 1 << 64)
    rng = np.random.default_rng(seed)
    out = []
    for low in rng.integers(0, 1 << 52, count, dtype=np.uint64).tolist():
        m = (home << 52) | low
        y = (m * cinv) & ((1 << 64) - 1)            # m = y * C  (mod 2^64)
        key = y ^ (y >> 32)                         # y = key ^ (key >> 32) is an involution
        assert (((key ^ (key >> 32)) * C) & ((1 << 64) - 1)) >> 52 == home
        out.append(key - (1 << 64) if key >= (1 << 63) else key)
    return np.array(out, dtype=np.int64)


@pytest.mark.parametrize("clash,path", [(4, "lowcard"), (14, "lowcard"), (40, "global")])
def test_lowcard_hash_mode_displaced_keys_and_overflow_list(pab, orc, clash, path):
    # `clash` keys share one home bucket of the packed table (they agree in the top 12 bits of lc_mix(key), the bucket is
    # the top 11): 4 fill the home bucket and the next one (two entries each), the next 16 go to the overflow list, more
    # than that and the kernel gives up in favour of the global-table path
    rng = np.random.default_rng(clash)
    n = 300_000
    pool = np.concatenate([_keys_with_home(1234, clash, clash), rng.integers(-2**62, 2**62, 300, dtype=np.int64)])
    frame = {"k": pa.array(pool[rng.integers(0, len(pool), n)]), "v": _rand_vals(rng, n)}
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", ["sum", "mean", "count", "first"], f"clash={clash} narrow")
    assert gb.timing()["path"] == path
    _cmp(gb, ora, rb, "v", ALL, f"clash={clash} all")
    assert gb.timing()["path"] == path


@pytest.mark.parametrize("G,nulls", [(2, False), (3, True), (17, False), (60, True), (130, False), (400, False)])
def test_lowcard_hash_mode_few_scattered_keys_replicated(pab, orc, G, nulls):
    # few SCATTERED keys (what hashed utf8 keys with a handful of values look like): the first pass on a handle runs
    # without accumulator replicas, every later one with 2^r replicas per id chosen from the now known group count
    # (lowcard.cuh, hash_rlog) — both against the oracle, both bit-identical to each other in everything but fp sums,
    # and each run-to-run deterministic
    rng = np.random.default_rng(G)
    n = 400_003
    pool = rng.integers(-2**62, 2**62, G, dtype=np.int64)
    k = pa.array(pool[rng.integers(0, G, n)], pa.int64(), mask=(rng.random(n) < 0.01) if nulls else None)
    frame = {"k": k, "v": _rand_vals(rng, n), "i": pa.array(rng.integers(-1000, 1000, n), pa.int64(), mask=rng.random(n) < 0.05)}
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", ALL, f"G={G} first pass")
    t1 = gb.timing()
    assert t1["path"] == "lowcard" and t1["mode"] == "hash" and t1["replication"] == 1
    r2 = _cmp(gb, ora, rb, "v", ALL, f"G={G} replicated pass")
    t2 = gb.timing()
    groups = G + (1 if nulls else 0)
    assert t2["path"] == "lowcard" and t2["mode"] == "hash" and t2["replication"] == (1 << min(5, int(np.log2(1000 // (groups + 1))))), t2
    _cmp(gb, ora, rb, "i", ALL, f"G={G} replicated pass, integers (double sum)")
    from util import assert_exact
    r3 = gb.aggregate(rb.column("v"), ALL)
    for a in ALL:
        assert_exact(r3[a], r2[a], f"run-to-run {a}")           # the replicated pass is deterministic too


def test_lowcard_dense_miss_reruns_in_hash_mode(pab, orc):
    # keys 0..499 everywhere except one far-away key in a row the 16384-row key sample does not see
    rng = np.random.default_rng(29)
    n = 1_000_000
    k = rng.integers(0, 500, n)
    k[777_777] = 10**15
    frame = {"k": pa.array(k, pa.int64()), "v": _rand_vals(rng, n)}
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", ALL, "dense miss")
    t = gb.timing()
    assert t["path"] == "lowcard" and t["mode"] == "hash" and t["passes"] == 2


@pytest.mark.parametrize("G,aggs,path", [(1000, ["sum", "mean", "count"], "lowcard"), (1001, ["sum", "count"], "global"),
                                         (1000, ["sum", "min", "max", "last"], "lowcard"), (1001, ["min", "max"], "global"),
                                         (2000, ["sum", "mean"], "global"), (5000, ALL, "global")])
def test_lowcard_capacity_boundaries(pab, orc, G, aggs, path):
    rng = np.random.default_rng(G)
    n = 300_000
    k = np.concatenate([np.arange(G), rng.integers(0, G, n - G)])      # every key present
    frame = {"k": pa.array(k * 3, pa.int64()), "v": _rand_vals(rng, n)}   # stride 3: a window wider than the table -> hash mode (1000 ids)
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", aggs, f"G={G} {aggs}")
    assert gb.timing()["path"] == path
    assert gb.groupSize() == G


@pytest.mark.parametrize("G,path", [(1024, "lowcard"), (1025, "global")])
def test_lowcard_dense_capacity_boundary(pab, orc, G, path):
    # dense (key - base) addressing holds 1024 ids; the hash-mode kernel 1000 (its packed table shares the 227 KB)
    rng = np.random.default_rng(G + 5)
    n = 300_000
    k = np.concatenate([np.arange(G), rng.integers(0, G, n - G)]) + 10**9
    frame = {"k": pa.array(k, pa.int64()), "v": _rand_vals(rng, n)}
    gb, ora, rb = _both(pab, orc, frame, "k")
    _cmp(gb, ora, rb, "v", ["sum", "mean", "count", "first"], f"dense G={G}")
    assert gb.timing()["path"] == path and gb.groupSize() == G


def test_int_values_mean_boundary_and_smem_front(pab, orc):
    # mean of integers needs the per-warp double sum: 640 groups on the shared-memory path, beyond that the
    # global path with its shared-memory front table (and, past its capacity, spills to the global table)
    rng = np.random.default_rng(41)
    n = 400_000
    for G, path, mode in [(640, "lowcard", "dense"), (700, "global", "smem-front"), (9000, "global", "smem-front")]:
        k = np.concatenate([np.arange(G), rng.integers(0, G, n - G)])
        frame = {"k": pa.array(k, pa.int64()), "v": pa.array(rng.integers(-1000, 1000, n), pa.int64(), mask=rng.random(n) < 0.05)}
        gb, ora, rb = _both(pab, orc, frame, "k")
        _cmp(gb, ora, rb, "v", ALL, f"int values G={G}")
        t = gb.timing()
        assert t["path"] == path and t["mode"] == mode, t


def test_partitioned_high_cardinality_equals_direct_scan(pab):
    # More groups than one shared-memory table holds: rows are radix-partitioned into buckets first (bucketed.cuh).
    # The oracle needs ~5 us per group and aggregate, so at this size the partitioned pass is checked against the
    # direct global-table scan (itself oracle-checked above at up to 250 K groups); tests/test_parity_large_gpu.py
    # holds the bucketed path to the oracle.
    import torch
    from util import assert_exact, assert_fp_close
    n, G = 6_000_001, 2_500_000
    k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
    pab.synth.keys(k, G); pab.synth.vals(v)
    k.mul_(7919).add_(-(10**9))                       # not dense, negative keys too
    torch.cuda.synchronize()
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    a = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G)
    b = pab.GroupBy("k", {"k": dk, "v": dv}, expected_groups=G, no_partition=True)
    ra, rb_ = a.aggregate(dv, ALL), b.aggregate(dv, ALL)
    assert a.timing()["mode"] == "bucketed" and b.timing()["mode"] is None
    assert a.groupSize() == b.groupSize() and a.unique().equals(b.unique())      # same groups, same first-appearance order
    for name in ALL:
        (assert_fp_close if name in ("sum", "mean") else assert_exact)(ra[name], rb_[name], name)
    assert sum(ra["count"].to_pylist()) == n


# ---------------- Grouper::Consume equivalent: per-row group ids ----------------
def _expected_ids(key_cols):
    seen, ids = {}, []
    for t in zip(*[c.to_pylist() for c in key_cols]):
        ids.append(seen.setdefault(t, len(seen)))
    return ids


@pytest.mark.parametrize("G,path", [(7, "auto"), (900, "auto"), (5000, "auto"), (70_000, "global")])
def test_row_ids_match_first_appearance_numbering(pab, G, path):
    rng = np.random.default_rng(G)
    n = 120_001
    k = pa.array(rng.integers(-G // 2, G // 2 + 1, n), pa.int64(), mask=rng.random(n) < 0.01)
    rb = pa.record_batch({"k": k, "v": pa.array(rng.random(n))})
    g = pab.GroupBy("k", rb, path=path)
    ids = g.row_ids()
    assert ids.type == pa.uint32() and len(ids) == n
    assert ids.to_pylist() == _expected_ids([k])
    # ids index unique(): gather the keys back
    assert g.unique().take(ids).equals(k)
    g.sum("v")                                           # ids stay valid after an aggregate on the same handle
    assert g.row_ids().equals(ids)


def test_row_ids_composite_int32_and_resample(pab):
    rng = np.random.default_rng(8)
    n = 50_000
    k1 = pa.array(rng.integers(0, 20, n).astype(np.int32), mask=rng.random(n) < 0.02)
    k2 = pa.array(np.array(["a", "bb", "ccc"])[rng.integers(0, 3, n)]).dictionary_encode()
    rb = pa.record_batch({"k1": k1, "k2": k2, "v": pa.array(rng.random(n))})
    g = pab.GroupBy(["k1", "k2"], rb)
    assert g.row_ids().to_pylist() == _expected_ids([k1, k2.indices])
    ts = np.sort(rng.integers(0, 3600 * 10**9, n)) + 1_600_000_000 * 10**9
    r = pab.resample({"v": rb.column("v")}, pa.array(ts, pa.timestamp("ns")), 60 * 10**9)
    ids = np.asarray(r.row_ids())
    labels = r.index().cast(pa.int64()).to_numpy()
    assert (labels[ids] == (ts // (60 * 10**9)) * (60 * 10**9)).all()


# ---------------- deferred (asynchronous) aggregate ----------------
@pytest.mark.parametrize("G,expect", [(1000, ("lowcard", "dense", 1)), (5000, ("global", "smem-front", 2)),
                                      (300_000, ("global", "smem-front", 2))])
def test_async_aggregate_equals_sync(pab, G, expect):
    import torch
    from util import assert_exact, assert_fp_close
    n = 2_000_000
    k = torch.empty(n, dtype=torch.int64, device="cuda"); v = torch.empty(n, dtype=torch.float64, device="cuda")
    pab.synth.keys(k, G); pab.synth.vals(v)
    torch.cuda.synchronize()
    dk, dv = pab.DeviceColumn.from_torch(k), pab.DeviceColumn.from_torch(v)
    a = pab.GroupBy("k", {"k": dk, "v": dv}); b = pab.GroupBy("k", {"k": dk, "v": dv})
    assert a.aggregate(dv, ALL, fetch=False, wait=False) == {}
    t = a.timing()                                        # completes the pass (falls back synchronously if needed)
    assert (t["path"], t["mode"], t["passes"]) == expect
    rb_ = b.aggregate(dv, ALL)
    assert a.unique().equals(b.unique())
    for name in ALL:
        (assert_fp_close if name in ("sum", "mean") else assert_exact)(a.fetch(name), rb_[name], name)
    # scattered keys: the dense kernel declines, the deferred pass is redone in hash mode
    k.mul_(-3335678366873096957)
    torch.cuda.synchronize()
    c = pab.GroupBy("k", {"k": dk, "v": dv})
    c.aggregate(dv, ["sum", "count"], fetch=False, wait=False)
    assert c.groupSize() == a.groupSize()
    if G == 1000:
        assert c.timing()["mode"] == "hash"
        assert c.fetch("count").equals(a.fetch("count"))


@pytest.mark.parametrize("groups", [7, 16, 256, 1000])
@pytest.mark.parametrize("nullable", [False, True])
def test_min_max_fp32_bounds_precheck(pab, orc, groups, nullable):
    """Dense-mode wide kernel, fp64 values: the {min, max} pre-check reads fp32 bounds (min rounded up, max rounded down)
    and only rows outside them reach the exact pair.  Values chosen to sit ON the rounding edges: doubles that no fp32
    holds repeated as a group's minimum / maximum, neighbours of such values one ulp away, magnitudes beyond fp32's range
    (round to +-inf / +-FLT_MAX), fp64 subnormals (round to +-0 / the smallest fp32 subnormal), signed zeros, +-inf, NaN,
    all-NaN groups.  Exact against the oracle, bit-identical (sign of zero included) against the hash-mode kernel — which
    keeps the plain exact pre-check — on the same values, and from run to run."""
    n = 600_000
    rng = np.random.default_rng(groups * 2 + nullable)
    k = rng.integers(0, groups, n)
    base = np.array([100.01, -100.01, 1.0 + 2.0**-30, 1.0 - 2.0**-31, 3.0e300, -3.0e300, 1.0e-310, -1.0e-310, 1.0e-45, -1.0e-45,
                     0.0, -0.0, np.inf, -np.inf, np.nan, 16777217.0, -16777217.0, 0.1, 0.3, 2.5])
    v = base[rng.integers(0, len(base), n)]
    # one ulp (of the double) around the edge values, so that rd / ru of neighbours land on the same fp32 pair
    step = rng.integers(-2, 3, n)
    fin = np.isfinite(v) & (v != 0)
    v[fin] = v[fin] * (1.0 + step[fin] * 2.0**-52)
    cont = rng.random(n) < 0.5
    v[cont] = rng.standard_normal(int(cont.sum())) * 1e3           # ordinary continuous values in between
    # which edge value is a group's extreme depends on the group: +-inf (k % 4 == 0), +-3e300 (1), +-16777217 (2), +-100.01 (3)
    c = k % 4
    v[(c >= 1) & np.isinf(v)] = 2.5
    v[(c >= 2) & (np.abs(v) > 1e299)] = 0.3
    v[(c == 3) & (np.abs(v) > 1.6e7)] = 0.1
    v[(c == 3) & cont] *= 1e-2
    v[k == 3] = np.nan                                              # a group with nothing but NaN
    v[k == 5] = np.where(rng.random(int((k == 5).sum())) < 0.5, 0.0, -0.0)   # nothing but zeros of both signs
    v[k == 6] = 100.01                                              # one value no fp32 holds, a million times
    mask = (rng.random(n) < 0.1) if nullable else None
    frame = {"k": pa.array(k, pa.int64()), "v": pa.array(v, pa.float64(), mask=mask)}
    aggs = ["sum", "min", "max", "count", "last"]
    gb, ora, rb = _both(pab, orc, frame, "k")
    res = _cmp(gb, ora, rb, "v", aggs, f"bounds pre-check G={groups}")
    assert gb.timing()["path"] == "lowcard" and gb.timing()["mode"] == "dense"
    res2 = gb.aggregate(frame["v"], aggs)
    # the same rows under scattered 64-bit keys: hash mode, exact pre-check only
    ks = (k.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D) + np.uint64(0x1234567)).astype(np.int64)
    gh = pab.GroupBy("k", pa.record_batch({"k": pa.array(ks, pa.int64()), "v": frame["v"]}))
    resh = gh.aggregate(frame["v"], aggs)
    assert gh.timing()["mode"] == "hash"
    order = {key: i for i, key in enumerate(gb.unique(0).to_pylist())}
    perm = pa.array([order[_unscramble(x)] for x in gh.unique(0).to_pylist()])
    for a in ("min", "max"):
        b1 = res[a].to_numpy(zero_copy_only=False).view(np.uint64)
        b2 = res2[a].to_numpy(zero_copy_only=False).view(np.uint64)
        assert np.array_equal(b1, b2), f"{a}: differs from run to run"
        bh = np.empty_like(b1)
        bh[np.asarray(perm)] = resh[a].to_numpy(zero_copy_only=False).view(np.uint64)
        nan = np.isnan(b1.view(np.float64))
        assert np.array_equal(b1[~nan], bh[~nan]), f"{a}: dense and hash mode differ (bits)"
        assert np.array_equal(np.asarray(res[a].is_valid()), np.asarray(res2[a].is_valid()))


def _unscramble(x):
    inv = pow(0x2545F4914F6CDD1D, -1, 2**64)
    return ((int(x) % 2**64 - 0x1234567) * inv) % 2**64
