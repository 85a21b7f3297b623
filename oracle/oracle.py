"""ctypes binding of oracle/_build/liboracle_pa.so (TEST INFRASTRUCTURE ONLY).

The oracle restates the reference's Arrow call sequence on the CPU (see oracle_groupby.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; nothing under pandasarrow_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import pyarrow as pa

from . import build_oracle

ORIGIN = {"epoch": 0, "start": 1, "start_day": 2, "end": 3, "end_day": 4, "custom": 5}


class _ArrowArray(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64),
                ("n_buffers", C.c_int64), ("n_children", C.c_int64), ("buffers", C.c_void_p),
                ("children", C.c_void_p), ("dictionary", C.c_void_p), ("release", C.c_void_p),
                ("private_data", C.c_void_p)]


class _ArrowSchema(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p),
                ("flags", C.c_int64), ("n_children", C.c_int64), ("children", C.c_void_p),
                ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.LIB
        if not os.path.exists(path):
            path = build_oracle.build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_groupby_create.restype = C.c_void_p
        L.orc_groupby_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_char_p), C.c_int, C.c_int]
        L.orc_groupby_destroy.argtypes = [C.c_void_p]
        L.orc_groupby_num_groups.restype = C.c_int64
        L.orc_groupby_num_groups.argtypes = [C.c_void_p]
        L.orc_groupby_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_groupby_unique.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_groupby_row_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_groupby_agg.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_groupby_min_max.argtypes = [C.c_void_p, C.c_char_p, C.c_int] + [C.c_void_p] * 4
        L.orc_groupby_group_slice.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_resample_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                          C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.orc_resample_labels_calendar.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_array_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_downsample_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_void_p]
        L.orc_scalar_agg.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def _export(obj):
    a, s = _ArrowArray(), _ArrowSchema()
    obj._export_to_c(C.addressof(a), C.addressof(s))
    return a, s


def _import_array(a, s) -> pa.Array:
    return pa.Array._import_from_c(C.addressof(a), C.addressof(s))


class OracleGroupBy:
    """pd::GroupBy as the reference computes it (group_by.h:22-247)."""

    def __init__(self, frame: pa.RecordBatch, keys, index: Optional[pa.Array] = None,
                 materialize: bool = False):
        if isinstance(keys, str):
            keys = [keys]
        L = lib()
        fa, fs = _export(frame)
        if index is not None:
            ia, isch = _export(index)
            ip, isp = C.addressof(ia), C.addressof(isch)
        else:
            ip = isp = None
        names = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
        self._h = L.orc_groupby_create(C.addressof(fa), C.addressof(fs), ip, isp, names, len(keys),
                                       int(materialize))
        if not self._h:
            raise OracleError(L.orc_last_error().decode())
        self.n_keys = len(keys)

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_groupby_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def num_groups(self) -> int:
        return lib().orc_groupby_num_groups(self._h)

    def timing_ms(self) -> dict:
        t = (C.c_double * 4)()
        lib().orc_groupby_timing(self._h, t)
        return {"consume": t[0], "groupings": t[1], "gather": t[2], "total": t[3]}

    def unique(self, key_i: int = 0) -> pa.Array:
        a, s = _ArrowArray(), _ArrowSchema()
        _check(lib().orc_groupby_unique(self._h, key_i, C.addressof(a), C.addressof(s)))
        return _import_array(a, s)

    def row_ids(self) -> pa.Array:
        a, s = _ArrowArray(), _ArrowSchema()
        _check(lib().orc_groupby_row_ids(self._h, C.addressof(a), C.addressof(s)))
        return _import_array(a, s)

    def agg(self, func: str, column: str, nthreads: int = 1, with_validity: bool = False):
        a, s = _ArrowArray(), _ArrowSchema()
        if with_validity:
            va, vs = _ArrowArray(), _ArrowSchema()
            _check(lib().orc_groupby_agg(self._h, func.encode(), column.encode(), nthreads,
                                         C.addressof(a), C.addressof(s), C.addressof(va), C.addressof(vs)))
            out = _import_array(a, s)
            valid = _import_array(va, vs) if va.release else None
            return out, valid
        _check(lib().orc_groupby_agg(self._h, func.encode(), column.encode(), nthreads,
                                     C.addressof(a), C.addressof(s), None, None))
        return _import_array(a, s)

    def min_max(self, column: str, nthreads: int = 1):
        a, s, b, t = _ArrowArray(), _ArrowSchema(), _ArrowArray(), _ArrowSchema()
        _check(lib().orc_groupby_min_max(self._h, column.encode(), nthreads, C.addressof(a), C.addressof(s),
                                         C.addressof(b), C.addressof(t)))
        return _import_array(a, s), _import_array(b, t)

    def group_slice(self, column: str, j: int) -> pa.Array:
        a, s = _ArrowArray(), _ArrowSchema()
        _check(lib().orc_groupby_group_slice(self._h, column.encode(), j, C.addressof(a), C.addressof(s)))
        return _import_array(a, s)


def resample_labels(index: pa.Array, freq_ns: int, closed_right=False, label_right=False,
                    origin="start_day", origin_custom_ns=0, offset_ns=0) -> pa.Array:
    """Per-row bucket labels = makeGroupInfo(...).downsample() (resample.cpp:202-295, resample.h:19-43)."""
    ia, isch = _export(index)
    a, s = _ArrowArray(), _ArrowSchema()
    _check(lib().orc_resample_labels(C.addressof(ia), C.addressof(isch), freq_ns, int(closed_right),
                                     int(label_right), ORIGIN[origin], origin_custom_ns, offset_ns,
                                     C.addressof(a), C.addressof(s)))
    return _import_array(a, s)


OFFSET_TYPES = {"D": 0, "M": 1, "QS": 2, "Q": 3, "WS": 4, "W": 5, "MS": 6, "Y": 7, "YS": 8}   # DateOffset::Type, core.h:122-134


def resample_labels_calendar(index: pa.Array, code: str, multiplier: int = 1, closed_right=True, label_right=False) -> pa.Array:
    """Per-row labels of pd::resample with a DateOffset rule (resample.cpp:248-267 + resample.h:19-43)."""
    ia, isch = _export(index)
    a, s = _ArrowArray(), _ArrowSchema()
    _check(lib().orc_resample_labels_calendar(C.addressof(ia), C.addressof(isch), OFFSET_TYPES[code], multiplier,
                                              int(closed_right), int(label_right), C.addressof(a), C.addressof(s)))
    return _import_array(a, s)


def resample_calendar(frame: pa.RecordBatch, index: pa.Array, code: str, multiplier: int = 1, **kw) -> "OracleGroupBy":
    labels = resample_labels_calendar(index, code, multiplier, **kw)
    return OracleGroupBy(frame, "__resampler_idx__", index=labels)


def array_sort(values: pa.Array, ascending: bool = True, take: bool = False) -> pa.Array:
    """Series::argsort (take=False: the uint64 indices) / Series::sort (take=True: the sorted values): arrow's
    array_sort_indices [+ Take] exactly as series.cpp:864-868,978-992 call them."""
    ia, isch = _export(values)
    a, s = _ArrowArray(), _ArrowSchema()
    _check(lib().orc_array_sort(C.addressof(ia), C.addressof(isch), int(ascending), int(take), C.addressof(a), C.addressof(s)))
    return _import_array(a, s)


def downsample_labels(index: pa.Array, multiple: int, unit: str, closed_label_right=False,
                      week_starts_monday=True, start_epoch=True) -> pa.Array:
    """DataFrame::downsample's per-row labels (dataframe.cpp:1265-1290)."""
    ia, isch = _export(index)
    a, s = _ArrowArray(), _ArrowSchema()
    _check(lib().orc_downsample_labels(C.addressof(ia), C.addressof(isch), multiple, unit.encode()[0:1],
                                       int(closed_label_right), int(week_starts_monday), int(start_epoch),
                                       C.addressof(a), C.addressof(s)))
    return _import_array(a, s)


def scalar_agg(array: pa.Array, func: str, skip_null: bool = True) -> pa.Scalar:
    """NDFrame<T>::sum/mean/min/max/count (ndframe.cpp:26-55,119)."""
    ia, isch = _export(array)
    a, s = _ArrowArray(), _ArrowSchema()
    _check(lib().orc_scalar_agg(C.addressof(ia), C.addressof(isch), func.encode(), int(skip_null),
                                C.addressof(a), C.addressof(s)))
    return _import_array(a, s)[0]


def resample(frame: pa.RecordBatch, index: pa.Array, freq_ns: int, **kw) -> OracleGroupBy:
    """pd::resample (resample.h:91-122): relabel the index, then group on it (group_by.h:258)."""
    labels = resample_labels(index, freq_ns, **kw)
    return OracleGroupBy(frame, "__resampler_idx__", index=labels)
