"""Build recipe for the oracle (TEST INFRASTRUCTURE ONLY — see oracle_groupby.cpp header).

Compiles oracle/oracle_groupby.cpp against the Arrow C++ headers and shared libraries that
ship inside the pyarrow wheel (Arrow 24.0.0) into oracle/_build/liboracle_pa.so.

The reference library itself is NOT compiled (oracle/_ref is therefore never produced): its
headers require Boost.date_time, oneTBB, range-v3, spdlog, tabulate and hosseinmoein/DataFrame,
none of which are installed, and its CMake needs files outside the tree.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_pa.so")
SRC = os.path.join(HERE, "oracle_groupby.cpp")


def _arrow_paths():
    import pyarrow

    inc = pyarrow.get_include()
    libdir = pyarrow.get_library_dirs()[0]
    libs = sorted(f for f in os.listdir(libdir) if f.startswith(("libarrow.so.", "libarrow_compute.so."))
                  and f.count(".") == 2)
    return inc, libdir, libs


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    inc, libdir, libs = _arrow_paths()
    cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wno-deprecated-declarations",
           "-I", inc, SRC, "-o", LIB, "-L", libdir, f"-Wl,-rpath,{libdir}"] + [f"-l:{l}" for l in libs]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
