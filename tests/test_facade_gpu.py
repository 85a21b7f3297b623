"""Runs the C++ façade tests (tests/cpp/facade_tests.cpp: the reference's own Catch2 cases restated
against pd::GroupBy / pd::Resampler / pd::resample over the C ABI).  Needs a GPU: -m gpu."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_facade_reference_cases():
    from pandasarrow_b200 import build_host
    _, exe = build_host.build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


def test_cpp_facade_builds_and_links():
    """CPU check: the façade and its test driver compile and link against libpa_b200.so."""
    from pandasarrow_b200 import build_host
    lib, exe = build_host.build()
    assert os.path.exists(lib) and os.path.exists(exe)
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for sym in ("GroupBy", "Resampler", "resample", "downsample"):
        assert sym in out
