"""Model of the fp32 bounds pre-check of the wide dense k_lowcard_scan (lowcard.cuh, lc_wide_update<VC_F, BOUNDS>):
per group a word {min rounded UP to fp32, max rounded DOWN to fp32}; a row is done after the word when
    (rd(v) > word.min or bits(rd(v)) == bits(word.min)) and (ru(v) < word.max or bits(ru(v)) == bits(word.max)).
The property the kernel relies on: a row that passes cannot change the exact {min, max} pair kept in the order map
(-0.0 < +0.0, NaN never a member) — for ordinary values, values on fp32 rounding edges, beyond fp32's range, fp64
subnormals and zeros of both signs.  Pure numpy; the CUDA path itself is checked by test_min_max_fp32_bounds_precheck."""
import numpy as np


def _rd(v):
    """double -> float32 rounded towards -inf (cvt.rm.f32.f64)"""
    with np.errstate(over="ignore"):
        f = v.astype(np.float32)
    up = f.astype(np.float64) > v
    with np.errstate(over="ignore"):
        f[up] = np.nextafter(f[up], np.float32(-np.inf))
    return f


def _ru(v):
    with np.errstate(over="ignore"):
        f = v.astype(np.float32)
    dn = f.astype(np.float64) < v
    with np.errstate(over="ignore"):
        f[dn] = np.nextafter(f[dn], np.float32(np.inf))
    return f


def _ord(v):
    """the order map of common.cuh (f64_to_ord): unsigned order == numeric order with -0.0 < +0.0"""
    b = v.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def _values(rng, n):
    edge = np.array([100.01, -100.01, 1.0 + 2.0**-30, 1.0 - 2.0**-31, 3.0e300, -3.0e300, 1e-310, -1e-310, 1e-45, -1e-45,
                     1.4e-45, -1.4e-45, 0.0, -0.0, np.inf, -np.inf, 16777217.0, -16777217.0, 3.4028234663852886e38,
                     3.4028235e38, -3.4028235e38, 0.1, 2.5, 5e-324, -5e-324])
    v = edge[rng.integers(0, len(edge), n)]
    wiggle = rng.integers(-2, 3, n)
    fin = np.isfinite(v) & (v != 0)
    v[fin] = v[fin] * (1.0 + wiggle[fin] * 2.0**-52)
    cont = rng.random(n) < 0.3
    v[cont] = rng.standard_normal(int(cont.sum())) * 10.0 ** rng.integers(-3, 4, int(cont.sum()))
    return v


def test_rounding_helpers_bracket_the_value():
    rng = np.random.default_rng(0)
    v = _values(rng, 200_000)
    lo, hi = _rd(v), _ru(v)
    assert (lo.astype(np.float64) <= v).all() and (hi.astype(np.float64) >= v).all()
    same = lo.view(np.uint32) == hi.view(np.uint32)
    assert (lo[same].astype(np.float64) == v[same]).all()                       # equal only when v is an fp32 value
    with np.errstate(over="ignore"):
        assert (np.nextafter(lo[~same], np.float32(np.inf)) == hi[~same]).all() # else adjacent fp32 values


def test_a_row_inside_the_word_cannot_change_the_exact_pair():
    rng = np.random.default_rng(1)
    n = 400_000
    v = _values(rng, n)                       # the row
    a = _values(rng, n)                       # a value already applied to the group: vouches for the min word
    b = _values(rng, n)                       # another one: vouches for the max word
    # exact pair as the kernel keeps it after a and b: min = the smaller, max = the larger in the order map
    lo_ab = np.where(_ord(a) <= _ord(b), a, b)
    hi_ab = np.where(_ord(a) >= _ord(b), a, b)
    w_min = _ru(lo_ab)                        # tightest words the kernel can hold for that pair
    w_max = _rd(hi_ab)
    lo, hi = _rd(v), _ru(v)
    in_lo = (lo > w_min) | (lo.view(np.uint32) == w_min.view(np.uint32))
    in_hi = (hi < w_max) | (hi.view(np.uint32) == w_max.view(np.uint32))
    skip = in_lo & in_hi
    assert skip.any() and (~skip).any()
    assert (_ord(v[skip]) >= _ord(lo_ab[skip])).all(), "a skipped row would have been a new minimum"
    assert (_ord(v[skip]) <= _ord(hi_ab[skip])).all(), "a skipped row would have been a new maximum"
    # looser (stale) words are only more conservative: every row that passes a looser word also passes nothing it should not
    looser_min = np.maximum(w_min, _ru(hi_ab))
    in_lo2 = (lo > looser_min) | (lo.view(np.uint32) == looser_min.view(np.uint32))
    assert (_ord(v[in_lo2]) >= _ord(lo_ab[in_lo2])).all()


def test_zero_of_the_other_sign_goes_on_to_the_exact_pair():
    z = np.array([0.0, -0.0, 0.0, -0.0])
    w = np.array([0.0, 0.0, -0.0, -0.0], dtype=np.float32)      # min word
    lo = _rd(z)
    in_lo = (lo > w) | (lo.view(np.uint32) == w.view(np.uint32))
    assert in_lo.tolist() == [True, False, False, True]
    # an unset word (NaN) passes nobody
    nanw = np.full(4, np.nan, dtype=np.float32)
    assert not ((lo > nanw) | (lo.view(np.uint32) == nanw.view(np.uint32))).any()
