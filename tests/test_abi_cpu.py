"""CPU-side checks: the C-ABI library loads and exports every symbol include/pa_b200.h declares,
computing calls fail loudly without a GPU, host generator self-checks.  No compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pandasarrow_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pa_b200.h")).read()
    declared = set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("pa_groupby")  # the struct typedef
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"declared in pa_b200.h but not exported: {missing}"
    assert set(_lib.EXPORTS) == declared


def test_struct_layouts_match_arrow_abi():
    from pandasarrow_b200 import _lib
    assert C.sizeof(_lib.ArrowArray) == 80 and C.sizeof(_lib.ArrowSchema) == 72
    assert C.sizeof(_lib.ArrowDeviceArray) == 80 + 8 + 8 + 8 + 24
    assert C.sizeof(_lib.PaOptions) == 4 + 4 + 8 + 8 + 8 + 32


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pyarrow as pa
    import pandasarrow_b200 as p
    with pytest.raises(p.PaError, match="no CUDA device"):
        p.GroupBy("k", pa.record_batch({"k": pa.array([1, 2, 1])}))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pandasarrow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inl")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "__none__", f"{f} mentions the oracle"


def test_host_generator_matches_survey_definition():
    from pandasarrow_b200 import hostgen as hg

    def sm(x):
        M = (1 << 64) - 1
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)

    n = 1000
    assert hg.keys(n, 1000).tolist() == [sm(i ^ 42) % 1000 for i in range(n)]
    assert hg.vals(n).tolist() == [(sm(i + 1337) >> 11) * 2.0 ** -53 for i in range(n)]
    assert hg.valid_mask(n).tolist() == [sm(i + 7) % 10 != 0 for i in range(n)]
    ts = hg.timestamps(n)
    assert (np.diff(ts) > 0).all() and ts[0] >= hg.T0_NS
    # shards concatenate to the whole (row-range sharding)
    assert np.array_equal(np.concatenate([hg.keys(400, 77, 0), hg.keys(600, 77, 400)]), hg.keys(1000, 77))


def test_bench_reference_arm_prints_contract_json():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port) runs without a GPU and prints one
    JSON line with the contract's keys (tiny sample here; the default sample is 100 M rows)."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-rows", "200000", "--groups", "100"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "groupby_agg_rows_per_s" and line["unit"] == "rows/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_header_is_plain_c_and_cxx():
    """The drop-in boundary is a C ABI: include/pa_b200.h must compile as C99 and as C++ on its own."""
    import subprocess, tempfile
    hdr = os.path.join(ROOT, "include", "pa_b200.h")
    for lang, std, comp in (("c", "-std=c99", "gcc"), ("c++", "-std=c++17", "g++")):
        with tempfile.NamedTemporaryFile("w", suffix=".c" if lang == "c" else ".cpp", delete=False) as f:
            f.write(f'#include "{hdr}"\nint main(void) {{ pa_options o; pa_options_init(&o); return (int)sizeof(struct ArrowDeviceArray) == 0; }}\n')
        r = subprocess.run([comp, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", lang, f.name], capture_output=True, text=True)
        os.unlink(f.name)
        assert r.returncode == 0, r.stderr
