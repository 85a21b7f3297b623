"""Python model of arrow::compute Floor/CeilTemporal (no time zone) to be ported to CUDA; checked against pyarrow."""
import numpy as np, pyarrow as pa, pyarrow.compute as pc

def fdiv(a, b): return a // b   # python floor division

def days_from_civil(y, m, d):
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468

def civil_from_days(z):
    z += 719468
    era = (z if z >= 0 else z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = mp + (3 if mp < 10 else -9)
    return (y + (m <= 2), m, d)

UNIT_NS = {'N': 1, 'U': 10**3, 'L': 10**6, 'S': 10**9, 'T': 60 * 10**9, 'H': 3600 * 10**9, 'D': 86400 * 10**9}
PARENT = {'N': 10**3, 'U': 10**6, 'L': 10**9, 'S': 60 * 10**9, 'T': 3600 * 10**9, 'H': 86400 * 10**9}
DAY = 86400 * 10**9

def trunc_div(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q

def floor_t(t, mult, unit, wsm, cbo):
    """t in ns since epoch."""
    if unit in UNIT_NS:
        u = UNIT_NS[unit]
        if mult == 1:
            return fdiv(t, u) * u
        if cbo:
            if unit == 'D':
                y, m, _ = civil_from_days(fdiv(t, DAY))
                origin = days_from_civil(y, m, 1) * DAY
            else:
                origin = fdiv(t, PARENT[unit]) * PARENT[unit]
            mm = mult * u
            return trunc_div(t - origin, mm) * mm + origin
        d = fdiv(t, u)
        m = (d // mult) * mult if d >= 0 else trunc_div(d - mult + 1, mult) * mult
        return m * u
    if unit == 'W':
        off = (3 if wsm else 4) * DAY
        tt = t + off
        W = 7 * DAY
        d = fdiv(tt, W)
        if mult == 1:
            return d * W - off
        if cbo:
            # start = last Thursday (Mon-start) / Wednesday (Sun-start) of December of the previous year + (mon - thu) days
            y, _, _ = civil_from_days(fdiv(tt, DAY))
            wd_target = 4 if wsm else 3          # thu = 4, wed = 3 (sun = 0)
            dec31 = days_from_civil(y - 1, 12, 31)
            wd31 = (dec31 + 4) % 7               # 1970-01-01 was a Thursday (4)
            last = dec31 - ((wd31 - wd_target) % 7)
            start = (last + 4) * DAY             # date::weekday difference (mon - thu) = 4 days: the Monday / Sunday after
            unitw = mult * W
            return trunc_div(tt - start, unitw) * unitw + start   # (arrow does not take the weekday offset off again here)
        m = (d // mult) * mult if d >= 0 else trunc_div(d - mult + 1, mult) * mult
        return m * W - off
    if unit in ('M', 'Q'):
        mul = mult * (3 if unit == 'Q' else 1)
        y, m, _ = civil_from_days(fdiv(t, DAY))
        if mul == 1:
            return days_from_civil(y, m, 1) * DAY
        if cbo:
            m0 = m - 1
            m0 -= m0 % mul
            return days_from_civil(y, m0 + 1, 1) * DAY
        tm = y * 12 + m - 1 - 1970 * 12
        tm = (tm // mul) * mul if tm >= 0 else trunc_div(tm - mul + 1, mul) * mul
        yy, mm = 1970 + tm // 12, tm % 12
        return days_from_civil(yy, mm + 1, 1) * DAY
    if unit == 'Y':
        y, _, _ = civil_from_days(fdiv(t, DAY))
        yy = trunc_div(y, mult) * mult
        return days_from_civil(yy, 1, 1) * DAY
    raise ValueError(unit)

def add_units(f, mult, unit):
    if unit in UNIT_NS: return f + mult * UNIT_NS[unit]
    if unit == 'W': return f + mult * 7 * DAY
    y, m, d = civil_from_days(fdiv(f, DAY))
    if unit in ('M', 'Q'):
        mul = mult * (3 if unit == 'Q' else 1)
        tm = y * 12 + (m - 1) + mul
        return days_from_civil(tm // 12, tm % 12 + 1, 1) * DAY
    return days_from_civil(y + mult, 1, 1) * DAY

def ceil_t(t, mult, unit, wsm, cbo):
    f = floor_t(t, mult, unit, wsm, cbo)
    if unit in ('M', 'Q', 'Y'):
        return add_units(f, mult, unit)          # arrow: always the NEXT boundary for month / quarter / year
    return add_units(f, mult, unit) if f < t else f

PAUNIT = {'N': 'nanosecond', 'U': 'microsecond', 'L': 'millisecond', 'S': 'second', 'T': 'minute', 'H': 'hour', 'D': 'day', 'W': 'week', 'M': 'month', 'Q': 'quarter', 'Y': 'year'}

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    ts = np.concatenate([rng.integers(-2 * 10**18, 4 * 10**18, 3000), rng.integers(1.5e18, 1.7e18, 3000).astype(np.int64),
                         np.array([0, -1, 1, 86400 * 10**9, -86400 * 10**9, 1577836800 * 10**9])]).astype(np.int64)
    arr = pa.array(ts, pa.timestamp('ns'))
    bad = 0
    for unit in 'NULSTHDWMQY':
        for mult in (1, 2, 3, 5, 7, 13):
            for wsm in (True, False):
                for cbo in (False, True):
                    for ceil in (False, True):
                        fn = pc.ceil_temporal if ceil else pc.floor_temporal
                        want = fn(arr, multiple=mult, unit=PAUNIT[unit], week_starts_monday=wsm, ceil_is_strictly_greater=False,
                                  calendar_based_origin=cbo).cast(pa.int64()).to_numpy()
                        got = np.array([(ceil_t if ceil else floor_t)(int(t), mult, unit, wsm, cbo) for t in ts], dtype=object)
                        got = np.array([g if -2**63 <= g < 2**63 else 0 for g in got], dtype=np.int64)
                        ne = (got != want)
                        if ne.any():
                            bad += 1
                            i = np.nonzero(ne)[0][0]
                            print(f"MISMATCH unit={unit} mult={mult} wsm={wsm} cbo={cbo} ceil={ceil}: {int(ne.sum())} rows; t={ts[i]} got={got[i]} want={want[i]}")
    print("bad combos:", bad)
