"""Generates tests/golden/*.json.  Run from the repo root: python tests/golden/generate.py

reference_vectors.json   known-answer vectors copied from the reference's own Catch2 tests (file:line cited per case);
                         the expected values are the literals those tests assert, not something computed here.
oracle_seeded.json       outputs of the ORACLE (oracle/oracle_groupby.cpp = the reference's arrow::compute call
                         sequence against Arrow 24.0.0) on small seeded inputs, inputs included, doubles stored as
                         C99 hex strings so that the fixture is bit exact.  It freezes the oracle: a drift (another
                         Arrow build, an edit of the oracle) shows up in tests/test_golden_cpu.py; the CUDA path is
                         checked against the same file in tests/test_zz_golden_gpu.py.
"""
import json
import os
import sys

import numpy as np
import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

ALL = ["sum", "mean", "count", "min", "max", "first", "last", "product", "variance", "stddev", "count_distinct"]


def enc(arr: pa.Array):
    """Arrow array -> JSON list; doubles as hex strings, nulls as None."""
    out = []
    for x in arr.to_pylist():
        if x is None:
            out.append(None)
        elif isinstance(x, float):
            out.append(float(x).hex())
        elif isinstance(x, bool):
            out.append(bool(x))
        else:
            out.append(int(x))
    return out


def reference_vectors():
    ids = ["allen", "victor", "hannah", "allen", "victor", "hannah", "allen", "victor", "hannah", "allen"]
    gender = ["male", "female", "male", "male", "female", "male", "male", "female", "male", "male"]
    return [
        {"source": "tests/cudf_examples/dataframe_resample_test.cpp:8-69,71-250",
         "frame": {"id": ids, "gender": gender, "age": [16, 10, 10, 20, 30, 40, 15, 25, 35, 45], "height": [9, 9, 9, 9, 9, 8, 8, 8, 8, 8]},
         "types": {"age": "int32", "height": "int32"},
         "cases": [
             {"key": "id", "unique": ["allen", "victor", "hannah"]},
             {"key": "gender", "unique": ["male", "female"],
              "expect": {"mean:age": [25.857142857142858, 21.666666666666668], "mean:height": [8.428571428571429, 8.666666666666666],
                         "min:age": [10, 10], "max:age": [45, 30], "min:height": [8, 8], "max:height": [9, 9],
                         "sum:age": [181, 65], "sum:height": [59, 26], "count:age": [7, 3]}}]},
        {"source": "tests/dataframe_iterator_test.cpp:11-76",
         "frame": {"a": [1, 1, 3, 1, 1, 1, 3, 8, 2, 2], "b": [10, 9, 8, 7, 6, 5, 4, 3, 2, 1]},
         "types": {"a": "int32", "b": "int32"},
         "cases": [{"key": "a", "unique": [1, 3, 8, 2], "expect": {"sum:a": [5, 6, 8, 4], "sum:b": [37, 12, 3, 3]}}]},
        {"source": "tests/cudf_examples/dataframe_resample_test.cpp:252-305",
         "frame": {"high": [11.1, 20.2, 21.0, 15, 20], "low": [9.1, 9.2, 10.0, 5, 10], "close": [10.1, 15.2, 20.0, 15, 15],
                   "open": [10, 20.2, 10.0, 15, 10], "volume": [100, 200, 210, 1, 2], "day": [1, 1, 2, 2, 5]},
         "types": {"high": "float32", "low": "float32", "close": "float32", "open": "float32", "volume": "uint64", "day": "int64"},
         "cases": [{"key": "day", "unique": [1, 2, 5],
                    "expect": {"first:open": [10, 10.0, 10], "last:close": [15.2, 15, 15], "max:high": [20.2, 21.0, 20],
                               "min:low": [9.1, 5, 10], "sum:volume": [300, 211, 2]}, "float32_expect": True}]},
        {"source": "tests/series_resample_test.cpp:12-70 (9 one-minute ticks from 2000-01-01, values 0..8, rule 3T)",
         "resample": {"start_ns": 946684800 * 10**9, "step_ns": 60 * 10**9, "n": 9, "freq_ns": 180 * 10**9},
         "cases": [{"closed_right": False, "label_right": False, "labels_min": [0, 3, 6], "sum": [3, 12, 21]},
                   {"closed_right": False, "label_right": True, "labels_min": [3, 6, 9], "sum": [3, 12, 21]},
                   {"closed_right": True, "label_right": True, "labels_min": [0, 3, 6, 9], "sum": [0, 6, 15, 15]}]},
    ]


def seeded_groupby():
    rng = np.random.default_rng(20241018)
    n, G = 2000, 37
    k = rng.integers(0, G, n) * 1009 - 17
    km = rng.random(n) < 0.02
    vm = rng.random(n) < 0.1
    vm[(k // 1009) % 9 == 4] = True                                    # all-null groups
    f = rng.normal(1.0, 2.0, n)
    f[rng.random(n) < 0.01] = np.nan
    p = np.exp(rng.uniform(-0.01, 0.01, n))
    i = rng.integers(-5, 6, n)
    cols = {"k": pa.array(k, pa.int64(), mask=km), "f": pa.array(f, pa.float64(), mask=vm), "p": pa.array(p, pa.float64(), mask=vm),
            "i": pa.array(i, pa.int32(), mask=vm)}
    rb = pa.record_batch(cols)
    g = orc.OracleGroupBy(rb, "k")
    out = {"inputs": {c: enc(a) for c, a in cols.items()}, "types": {"k": "int64", "f": "float64", "p": "float64", "i": "int32"},
           "unique": enc(g.unique()), "results": {}}
    for col in ("f", "p", "i"):
        for a in ALL:
            if a in ("mean", "variance", "stddev"):
                vals, valid = g.agg(a, col, with_validity=True)
                out["results"][f"{a}:{col}"] = [v if ok else None for v, ok in zip(enc(vals), valid.to_pylist())]
            else:
                out["results"][f"{a}:{col}"] = enc(g.agg(a, col))
    return out


def seeded_resample():
    rng = np.random.default_rng(7)
    n = 1500
    ts = np.cumsum(rng.integers(1, 40_000_000_000, n)).astype(np.int64) + 1_577_836_800_000_000_000 + 3 * 3600 * 10**9
    v = rng.normal(size=n)
    vm = rng.random(n) < 0.05
    idx = pa.array(ts, pa.timestamp("ns"))
    val = pa.array(v, pa.float64(), mask=vm)
    out = {"ts": [int(x) for x in ts], "v": enc(val), "cases": []}
    for freq_ns, closed_right, label_right in ((60 * 10**9, False, False), (7 * 60 * 10**9, True, True), (3600 * 10**9, False, True)):
        labels = orc.resample_labels(idx, freq_ns, closed_right=closed_right, label_right=label_right)
        g = orc.OracleGroupBy(pa.record_batch({"k": labels, "v": val}), "k")
        case = {"freq_ns": freq_ns, "closed_right": closed_right, "label_right": label_right,
                "labels": [int(x) for x in g.unique().cast(pa.int64()).to_pylist()], "results": {}}
        for a in ("sum", "count", "min", "max", "first", "last"):
            case["results"][a] = enc(g.agg(a, "v"))
        vals, valid = g.agg("mean", "v", with_validity=True)
        case["results"]["mean"] = [x if ok else None for x, ok in zip(enc(vals), valid.to_pylist())]
        out["cases"].append(case)
    return out


if __name__ == "__main__":
    json.dump(reference_vectors(), open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)
    json.dump({"groupby": seeded_groupby(), "resample": seeded_resample()}, open(os.path.join(HERE, "oracle_seeded.json"), "w"))
    print("written", os.listdir(HERE))
