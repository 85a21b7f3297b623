"""pandasarrow_b200 — B200-native group-by hash aggregation behind PandasArrow's GroupBy/Resampler API.

The product is the CUDA library `lib/libpa_b200.so` (C ABI in include/pa_b200.h) plus the C++
façade in csrc/host/.  This Python package is a thin ctypes harness over the C ABI used by the
tests and bench.py; importing it loads the CUDA library and raises ImportError if it is missing
(there is no CPU fallback).
"""
from . import _lib

_lib.load()

from .groupby import DeviceColumn, GroupBy, MergedGroupBy, PaError, Resampler, Sorted, aggregate_chunked, to_device, downsample, resample, resample_calendar, synth  # noqa: E402

__all__ = ["DeviceColumn", "GroupBy", "MergedGroupBy", "Resampler", "Sorted", "aggregate_chunked", "to_device", "resample", "resample_calendar", "downsample", "PaError", "synth"]
