"""The CUDA path (through the C ABI) against the committed golden fixtures of tests/golden/ — the reference's own
known-answer vectors and the oracle's frozen outputs on seeded inputs.  Integer / positional results bit exact,
floating-point sums / means / products / variances within 1e-12 of the group's scale.  Needs a GPU: -m gpu.
(The comparison logic is exercised on CPU by tests/test_golden_cpu.py::test_golden_comparison_logic_selfcheck.)"""
import numpy as np
import pyarrow as pa
import pytest

import golden_util as gu

BASE = ["sum", "mean", "count", "min", "max", "first", "last"]
STAGE2 = ["product", "variance", "stddev"]


def compare_groupby_to_golden(d, unique_got, fetch):
    """d: the "groupby" fixture.  unique_got: our unique keys (any group order).  fetch(agg, col) -> pa.Array in OUR order."""
    from util import align_to, assert_exact, assert_fp_close
    perm = pa.array(align_to([(x,) for x in unique_got], [(x,) for x in d["unique"]]))
    keys = gu.dec(d["inputs"]["k"], pa.int64())
    for col in ("f", "p", "i"):
        in_type = gu.PA_TYPES[d["types"][col]]
        vals = gu.dec(d["inputs"][col], in_type)
        s1, m2 = gu.per_group_abs_scales(keys, vals.cast(pa.float64()), d["unique"])
        for agg in BASE + STAGE2 + ["count_distinct"]:
            want = gu.dec(d["results"][f"{agg}:{col}"], gu.result_type(agg, in_type))
            got = fetch(agg, col).take(perm)
            what = f"golden {agg}({col})"
            fp = pa.types.is_floating(want.type)
            if agg == "sum" and fp:
                assert_fp_close(got, want, what, s1)
            elif agg == "mean":
                assert_fp_close(got, want, what, np.sqrt(m2))
            elif agg == "variance":
                assert_fp_close(got, want, what, m2)
            elif agg == "stddev":
                assert_fp_close(got, want, what, np.sqrt(m2))
            elif agg == "product" and fp:
                assert_fp_close(got, want, what)
            else:
                assert_exact(got, want, what)


def compare_resample_to_golden(case, labels_got, fetch):
    from util import assert_exact, assert_fp_close
    assert labels_got == case["labels"], "bucket labels (time order)"
    for a in ("sum", "mean", "count", "min", "max", "first", "last"):
        want = gu.dec(case["results"][a], pa.int64() if a == "count" else pa.float64())
        (assert_fp_close if a in ("sum", "mean") else assert_exact)(fetch(a), want, f"golden resample {a} freq={case['freq_ns']}")


@pytest.mark.gpu
def test_cuda_path_matches_reference_vectors():
    import pandasarrow_b200 as pab
    for case in gu.load("reference_vectors.json"):
        if "resample" in case:
            r = case["resample"]
            ts = pa.array([r["start_ns"] + i * r["step_ns"] for i in range(r["n"])], pa.timestamp("ns"))
            vals = pa.array(range(r["n"]), pa.int64())
            for c in case["cases"]:
                rs = pab.resample({"v": vals}, ts, r["freq_ns"], closed_right=c["closed_right"], label_right=c["label_right"])
                got = [(x - r["start_ns"]) // (60 * 10**9) for x in rs.index().cast(pa.int64()).to_pylist()]
                assert got == c["labels_min"], case["source"]
                assert rs.aggregate(vals, ["sum"])["sum"].to_pylist() == c["sum"], case["source"]
            continue
        rb = gu.reference_frame(case)
        for c in case["cases"]:
            g = pab.GroupBy(c["key"], rb)
            assert g.unique().to_pylist() == c["unique"], case["source"]
            for what, want in c.get("expect", {}).items():
                agg, col = what.split(":")
                got = g.aggregate(rb.column(col), [agg])[agg].to_pylist()
                if c.get("float32_expect") and pa.types.is_floating(rb.column(col).type):
                    want = [float(np.float32(x)) for x in want]
                assert got == want, (case["source"], what)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["auto", "global"])
def test_cuda_path_matches_frozen_oracle_outputs(path):
    import pandasarrow_b200 as pab
    d = gu.load("oracle_seeded.json")["groupby"]
    cols = {c: gu.dec(v, gu.PA_TYPES[d["types"][c]]) for c, v in d["inputs"].items()}
    rb = pa.record_batch(cols)
    gb = pab.GroupBy("k", rb, path=path)
    cache = {}

    def fetch(agg, col):
        if (agg, col) not in cache:
            for group in (BASE, STAGE2, ["count_distinct"]):      # three calls per column, as the other tests make them
                cache.update({(a, col): v for a, v in gb.aggregate(rb.column(col), group).items()})
        return cache[(agg, col)]

    compare_groupby_to_golden(d, gb.unique().to_pylist(), fetch)
    r = gu.load("oracle_seeded.json")["resample"]
    idx, val = pa.array(r["ts"], pa.timestamp("ns")), gu.dec(r["v"], pa.float64())
    for c in r["cases"]:
        rs = pab.resample({"v": val}, idx, c["freq_ns"], closed_right=c["closed_right"], label_right=c["label_right"])
        res = rs.aggregate(val, ["sum", "mean", "count", "min", "max", "first", "last"])
        compare_resample_to_golden(c, rs.index().cast(pa.int64()).to_pylist(), lambda a: res[a])
