#!/bin/bash
# GPU batch 54: the clean rebuild of the last commit: smoke() + the tests of this session's kernel changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_resample_gpu.py tests/test_groupby_gpu.py -m gpu -q -x -k "special_values or bounds or config4 or nullable or golden" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_parity_large_gpu.py -m gpu -q -x -k "nullable_values_vs_oracle" 2>&1 | tail -2
