"""Shared helpers for the parity tests: run the CUDA path (through the C ABI) and the oracle on
the same inputs and compare with the bar BASELINE.json states (bit-exact for keys, counts,
min/max, integer sums, first/last; <= 1e-12 relative for fp64 sum/mean)."""
import math

import numpy as np
import pyarrow as pa

FP_RTOL = 1e-12   # north_star: "within 1e-12 relative error for fp64 sum/mean"


def with_abs(frame):
    """Adds |x| twins of the floating columns: sums of mixed-sign data are compared relative to
    sum(|x|) of the group (the conditioning of the problem), which equals |sum| for the
    non-negative benchmark data."""
    import pyarrow.compute as pc
    cols = dict(frame) if isinstance(frame, dict) else {n: frame.column(n) for n in frame.schema.names}
    out = dict(cols)
    for n, c in cols.items():
        if isinstance(c, (pa.Array, pa.ChunkedArray)) and pa.types.is_floating(c.type):
            out["__abs_" + n] = pc.abs(c).cast(pa.float64())
    return pa.record_batch(out)


def abs_scale(ora, column, mean=False):
    """Per-group sum(|x|) (or mean(|x|)) from the oracle, NaN/inf-free, or None."""
    try:
        a = ora.agg("mean" if mean else "sum", "__abs_" + column, nthreads=8)
    except Exception:
        return None
    s = a.to_numpy(zero_copy_only=False).astype(np.float64)
    return np.where(np.isfinite(s), s, 0.0)


def assert_fp_close(got: pa.Array, want: pa.Array, what="", scale=None):
    assert got.type == want.type, f"{what}: dtype {got.type} != {want.type}"
    assert len(got) == len(want), f"{what}: length {len(got)} != {len(want)}"
    gv = np.asarray(got.is_valid()); wv = np.asarray(want.is_valid())
    assert (gv == wv).all(), f"{what}: validity differs"
    g = got.to_numpy(zero_copy_only=False).astype(np.float64)[gv]
    w = want.to_numpy(zero_copy_only=False).astype(np.float64)[wv]
    sc = None if scale is None else np.asarray(scale, dtype=np.float64)[wv]
    nan = np.isnan(w)
    assert (np.isnan(g) == nan).all(), f"{what}: NaN pattern differs"
    g, w = g[~nan], w[~nan]
    sc = None if sc is None else sc[~nan]
    fin = np.isfinite(w)
    assert (g[~fin] == w[~fin]).all(), f"{what}: infinities differ"
    g, w = g[fin], w[fin]
    sc = None if sc is None else sc[fin]
    if len(w):
        denom = np.maximum(np.abs(w), np.finfo(np.float64).tiny)
        if sc is not None:
            denom = np.maximum(denom, sc)
        rel = np.abs(g - w) / denom
        assert rel.max() <= FP_RTOL, f"{what}: max rel err {rel.max():.3e} > {FP_RTOL}"


def assert_exact(got: pa.Array, want: pa.Array, what=""):
    assert got.type == want.type, f"{what}: dtype {got.type} != {want.type}"
    if pa.types.is_floating(got.type):
        # bit-exact except NaN payload / sign of zero (documented: Arrow's fmin/fmax tie rule)
        gv = np.asarray(got.is_valid()); wv = np.asarray(want.is_valid())
        assert (gv == wv).all(), f"{what}: validity differs"
        g = got.to_numpy(zero_copy_only=False)[gv]; w = want.to_numpy(zero_copy_only=False)[wv]
        assert ((g == w) | (np.isnan(g) & np.isnan(w))).all(), f"{what}: values differ"
    else:
        assert got.equals(want), f"{what}: {got.to_pylist()[:8]} != {want.to_pylist()[:8]}"


def first_appearance_order(key_cols):
    """Strict first-appearance order of the distinct key tuples (None = null)."""
    seen, out = set(), []
    cols = [c.to_pylist() for c in key_cols]
    for t in zip(*cols):
        if t not in seen:
            seen.add(t)
            out.append(t)
    return out


def align_to(ours_keys, oracle_keys):
    """Permutation p such that ours[p[j]] is the oracle's j-th group."""
    pos = {t: i for i, t in enumerate(ours_keys)}
    assert len(pos) == len(ours_keys), "duplicate groups in the CUDA result"
    assert set(pos) == set(oracle_keys), "key sets differ"
    return np.array([pos[t] for t in oracle_keys], dtype=np.int64)


def compare_all(gb, ora, frame, column, aggs, what="", key_cols=None):
    """gb: pandasarrow_b200.GroupBy, ora: oracle.OracleGroupBy built on the same frame/key.

    Group ORDER: the CUDA path emits strict first-appearance order (checked here against a plain
    Python scan of the key columns).  The reference inherits arrow::compute::Grouper's id order,
    which equals first appearance on every vector the reference tests hold but may swap keys that
    first appear inside the same internal mini-batch (unspecified, hash-collision dependent).
    Aggregates are therefore compared per key, after aligning the two orders."""
    col = frame[column] if isinstance(frame, dict) else frame.column(column)
    res = gb.aggregate(col, aggs)
    nk = ora.n_keys
    ours = list(zip(*[gb.unique(i).to_pylist() for i in range(nk)]))
    theirs = list(zip(*[ora.unique(i).to_pylist() for i in range(nk)]))
    for i in range(nk):
        assert gb.unique(i).type == ora.unique(i).type, f"{what}: key {i} dtype"
    if key_cols is not None:
        assert ours == first_appearance_order(key_cols), f"{what}: not in first-appearance order"
    perm = align_to(ours, theirs)
    is_float = pa.types.is_floating(col.type)
    for a in aggs:
        got = res[a].take(pa.array(perm))
        if a == "mean":
            want, valid = ora.agg("mean", column, nthreads=8, with_validity=True)
            # reference quirk (pd_core_macros.h:67): validity dropped; the C ABI keeps it, compare both views
            want = pa.array(want.to_numpy(zero_copy_only=False), pa.float64(),
                            mask=~np.asarray(valid.to_numpy(zero_copy_only=False), dtype=bool))
            assert_fp_close(got, want, f"{what} mean", abs_scale(ora, column, mean=True))
        elif a == "sum" and is_float:
            assert_fp_close(got, ora.agg("sum", column, nthreads=8), f"{what} sum", abs_scale(ora, column))
        else:
            assert_exact(got, ora.agg(a, column, nthreads=8), f"{what} {a}")
    return res
