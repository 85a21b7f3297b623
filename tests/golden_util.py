"""Loader for tests/golden/*.json (see tests/golden/generate.py for what the files hold)."""
import json
import os

import numpy as np
import pyarrow as pa

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PA_TYPES = {"int32": pa.int32(), "int64": pa.int64(), "uint64": pa.uint64(), "float32": pa.float32(), "float64": pa.float64()}


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def dec(values, typ: pa.DataType) -> pa.Array:
    """JSON list (doubles as C99 hex strings, nulls as None) -> Arrow array of `typ`."""
    if pa.types.is_floating(typ):
        vals = [None if x is None else (float.fromhex(x) if isinstance(x, str) else float(x)) for x in values]
    else:
        vals = values
    return pa.array(vals, typ)


def result_type(agg: str, in_type: pa.DataType) -> pa.DataType:
    """Result dtype of arrow::compute's scalar aggregates as the reference calls them."""
    if agg in ("mean", "variance", "stddev"):
        return pa.float64()
    if agg in ("count", "count_distinct"):
        return pa.int64()
    if agg in ("sum", "product"):
        if pa.types.is_floating(in_type):
            return pa.float64()
        return pa.uint64() if pa.types.is_unsigned_integer(in_type) else pa.int64()
    return in_type


def reference_frame(case):
    cols = {}
    for name, vals in case["frame"].items():
        t = case.get("types", {}).get(name)
        cols[name] = pa.array(vals, PA_TYPES[t]) if t else pa.array(vals)
    return pa.record_batch(cols)


def per_group_abs_scales(keys: pa.Array, vals: pa.Array, unique: list):
    """sum|x|, mean(x^2) per group (finite values only), in the order of `unique`."""
    pos = {k: i for i, k in enumerate(unique)}
    s1 = np.zeros(len(unique)); s2 = np.zeros(len(unique)); n = np.zeros(len(unique))
    for k, v in zip(keys.to_pylist(), vals.to_pylist()):
        if v is None or not np.isfinite(v):
            continue
        j = pos[k]
        s1[j] += abs(v); s2[j] += v * v; n[j] += 1
    return s1, s2 / np.maximum(n, 1)
