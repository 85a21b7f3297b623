"""Pins tests/temporal_model.py — the integer model that csrc/temporal.cuh transcribes — against pyarrow's own
floor_temporal / ceil_temporal kernels (the calls DataFrame::downsample makes, /root/reference/src/dataframe.cpp:1265-1290)
for every unit, several multiples, both week starts, both origins.  CPU only."""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from temporal_model import PAUNIT, ceil_t, floor_t


@pytest.mark.parametrize("unit", list("NULSTHDWMQY"))
def test_model_matches_arrow(unit):
    rng = np.random.default_rng(ord(unit))
    ts = np.concatenate([rng.integers(-2 * 10**18, 4 * 10**18, 300), rng.integers(15 * 10**17, 17 * 10**17, 300),
                         np.array([0, -1, 1, 86400 * 10**9, -86400 * 10**9, 1577836800 * 10**9])]).astype(np.int64)
    arr = pa.array(ts, pa.timestamp("ns"))
    for mult in (1, 2, 3, 7, 13):
        for wsm in (True, False):
            for cbo in (False, True):
                for ceil in (False, True):
                    fn = pc.ceil_temporal if ceil else pc.floor_temporal
                    want = fn(arr, multiple=mult, unit=PAUNIT[unit], week_starts_monday=wsm, ceil_is_strictly_greater=False,
                              calendar_based_origin=cbo).cast(pa.int64()).to_numpy()
                    got = np.array([(ceil_t if ceil else floor_t)(int(t), mult, unit, wsm, cbo) for t in ts], dtype=np.int64)
                    assert np.array_equal(got, want), (unit, mult, wsm, cbo, ceil)
