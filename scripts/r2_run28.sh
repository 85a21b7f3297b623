#!/bin/bash
# GPU batch 28: replicated hash mode: parity, then timing of the second pass for few scattered keys
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_groupby_gpu.py tests/test_strkeys_gpu.py tests/test_zz_golden_gpu.py -m gpu -q -x > gpurun_out/r2_pytest28.log 2>&1
tail -8 gpurun_out/r2_pytest28.log
for G in 2 8 32 128 1000; do
echo "== $G random 64-bit keys"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups $G --hashed --iters 3 2>&1 | grep "iter [02]" | cut -c1-100
done
