"""Whole-column aggregates of the C++ façade on the GPU, against arrow::compute's scalar kernels:
runs tests/cpp/scalar_agg_tests.cpp.  Needs a GPU: -m gpu.  (Sorted last on purpose: first verified on the GPU box at
the end of round 1.)"""
import subprocess

import pytest


def test_scalar_tests_build_and_link():
    from pandasarrow_b200 import build_host
    import os
    assert os.path.exists(build_host.build_scalar_tests())


@pytest.mark.gpu
def test_cpp_facade_whole_column_aggregates():
    from pandasarrow_b200 import build_host
    exe = build_host.build_scalar_tests()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
