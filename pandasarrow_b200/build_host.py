"""Builds the host-side C++ façade (csrc/host/pd_groupby.cpp -> lib/libpd_b200.so) and the C++ test
driver (tests/cpp/facade_tests) against pyarrow's Arrow C++ and libpa_b200.so."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_DIR = os.path.join(HERE, "lib")
FACADE = os.path.join(LIB_DIR, "libpd_b200.so")
TEST_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "facade_tests")
SRC = os.path.join(HERE, "csrc", "host", "pd_groupby.cpp")
HDR = os.path.join(HERE, "csrc", "host", "pd_groupby.h")
TEST_SRC = os.path.join(ROOT, "tests", "cpp", "facade_tests.cpp")
SCALAR_TEST_BIN = os.path.join(ROOT, "tests", "cpp", "_build", "scalar_agg_tests")
SCALAR_TEST_SRC = os.path.join(ROOT, "tests", "cpp", "scalar_agg_tests.cpp")


def _arrow():
    import pyarrow
    inc = pyarrow.get_include()
    libdir = pyarrow.get_library_dirs()[0]
    libs = sorted(f for f in os.listdir(libdir) if f.startswith(("libarrow.so.", "libarrow_compute.so.", "libparquet.so.")) and f.count(".") == 2)
    return inc, libdir, [f"-l:{l}" for l in libs]


def _stale(out, srcs):
    return not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)


def build(force: bool = False, verbose: bool = False):
    inc, libdir, libs = _arrow()
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(os.path.dirname(TEST_BIN), exist_ok=True)
    common = ["g++", "-std=c++20", "-O2", "-fPIC", "-Wall", "-Wno-deprecated-declarations", "-I", inc]
    if force or _stale(FACADE, [SRC, HDR]):
        cmd = common + ["-shared", SRC, "-o", FACADE, "-L", libdir, "-L", LIB_DIR, "-lpa_b200",
                        f"-Wl,-rpath,{libdir}", "-Wl,-rpath,$ORIGIN"] + libs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    if force or _stale(TEST_BIN, [TEST_SRC, HDR, FACADE]):
        cmd = common + [TEST_SRC, "-o", TEST_BIN, "-L", LIB_DIR, "-lpd_b200", "-lpa_b200", "-L", libdir,
                        f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{LIB_DIR}"] + libs
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return FACADE, TEST_BIN


def build_scalar_tests(force: bool = False):
    """tests/cpp/scalar_agg_tests (whole-column aggregates); kept apart from build() so that it cannot affect it."""
    inc, libdir, libs = _arrow()
    build(force)
    if force or _stale(SCALAR_TEST_BIN, [SCALAR_TEST_SRC, HDR, FACADE]):
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-Wall", "-Wno-deprecated-declarations", "-I", inc, SCALAR_TEST_SRC, "-o", SCALAR_TEST_BIN,
               "-L", LIB_DIR, "-lpd_b200", "-lpa_b200", "-L", libdir, f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{LIB_DIR}"] + libs
        subprocess.run(cmd, check=True)
    return SCALAR_TEST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_scalar_tests(force="--force" in sys.argv))
