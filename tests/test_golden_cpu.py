"""The committed golden fixtures (tests/golden/, made by tests/golden/generate.py) against the ORACLE — CPU only.
reference_vectors.json: the literals the reference's own tests assert.  oracle_seeded.json: the oracle's frozen
outputs on seeded inputs (bit exact) — a drift of the oracle (other Arrow build, an edit) fails here."""
import numpy as np
import pyarrow as pa

import golden_util as gu
from oracle import oracle as orc


def test_reference_vectors_hold_for_the_oracle():
    for case in gu.load("reference_vectors.json"):
        if "resample" in case:
            r = case["resample"]
            ts = pa.array([r["start_ns"] + i * r["step_ns"] for i in range(r["n"])], pa.timestamp("ns"))
            vals = pa.array(range(r["n"]), pa.int64())
            for c in case["cases"]:
                labels = orc.resample_labels(ts, r["freq_ns"], closed_right=c["closed_right"], label_right=c["label_right"])
                g = orc.OracleGroupBy(pa.record_batch({"k": labels, "v": vals}), "k")
                got = [(x - r["start_ns"]) // (60 * 10**9) for x in g.unique().cast(pa.int64()).to_pylist()]
                assert got == c["labels_min"], case["source"]
                assert g.agg("sum", "v").to_pylist() == c["sum"], case["source"]
            continue
        rb = gu.reference_frame(case)
        for c in case["cases"]:
            g = orc.OracleGroupBy(rb, c["key"])
            assert g.unique().to_pylist() == c["unique"], case["source"]
            for what, want in c.get("expect", {}).items():
                agg, col = what.split(":")
                got = g.agg(agg, col).to_pylist()
                if c.get("float32_expect") and pa.types.is_floating(rb.column(col).type):
                    want = [float(np.float32(x)) for x in want]
                assert got == want, (case["source"], what)


def test_oracle_reproduces_its_frozen_outputs_bit_for_bit():
    d = gu.load("oracle_seeded.json")["groupby"]
    cols = {c: gu.dec(v, gu.PA_TYPES[d["types"][c]]) for c, v in d["inputs"].items()}
    g = orc.OracleGroupBy(pa.record_batch(cols), "k")
    assert g.unique().to_pylist() == d["unique"]
    for what, want in d["results"].items():
        agg, col = what.split(":")
        if agg in ("mean", "variance", "stddev"):
            vals, valid = g.agg(agg, col, with_validity=True)
            got = [None if not ok else float(v).hex() for v, ok in zip(vals.to_pylist(), valid.to_pylist())]
        else:
            got = [None if v is None else (float(v).hex() if isinstance(v, float) else v) for v in g.agg(agg, col).to_pylist()]
        assert got == want, what
    r = gu.load("oracle_seeded.json")["resample"]
    idx = pa.array(r["ts"], pa.timestamp("ns"))
    val = gu.dec(r["v"], pa.float64())
    for c in r["cases"]:
        labels = orc.resample_labels(idx, c["freq_ns"], closed_right=c["closed_right"], label_right=c["label_right"])
        g = orc.OracleGroupBy(pa.record_batch({"k": labels, "v": val}), "k")
        assert g.unique().cast(pa.int64()).to_pylist() == c["labels"]
        for a in ("sum", "count", "min", "max", "first", "last"):
            got = [None if v is None else (float(v).hex() if isinstance(v, float) else v) for v in g.agg(a, "v").to_pylist()]
            assert got == c["results"][a], (c["freq_ns"], a)


def test_golden_comparison_logic_selfcheck():
    """Runs the comparison code of tests/test_zz_golden_gpu.py with the oracle standing in for the CUDA path (in a
    shuffled group order), so that a mistake in the harness itself cannot hide behind the GPU marker."""
    import test_zz_golden_gpu as zz
    d = gu.load("oracle_seeded.json")["groupby"]
    cols = {c: gu.dec(v, gu.PA_TYPES[d["types"][c]]) for c, v in d["inputs"].items()}
    g = orc.OracleGroupBy(pa.record_batch(cols), "k")
    order = np.random.default_rng(1).permutation(g.num_groups)
    shuffled = [d["unique"][i] for i in order]

    def fetch(agg, col):
        if agg in ("mean", "variance", "stddev"):
            vals, valid = g.agg(agg, col, with_validity=True)
            a = pa.array(vals.to_numpy(zero_copy_only=False), pa.float64(), mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool))
        else:
            a = g.agg(agg, col)
        return a.take(pa.array(order))

    zz.compare_groupby_to_golden(d, shuffled, fetch)
    r = gu.load("oracle_seeded.json")["resample"]
    idx, val = pa.array(r["ts"], pa.timestamp("ns")), gu.dec(r["v"], pa.float64())
    for c in r["cases"]:
        labels = orc.resample_labels(idx, c["freq_ns"], closed_right=c["closed_right"], label_right=c["label_right"])
        go = orc.OracleGroupBy(pa.record_batch({"k": labels, "v": val}), "k")

        def fetch_r(a):
            if a == "mean":
                vals, valid = go.agg("mean", "v", with_validity=True)
                return pa.array(vals.to_numpy(zero_copy_only=False), pa.float64(), mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool))
            return go.agg(a, "v")

        zz.compare_resample_to_golden(c, go.unique().cast(pa.int64()).to_pylist(), fetch_r)
