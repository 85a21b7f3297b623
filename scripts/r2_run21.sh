#!/bin/bash
# GPU batch 21: hash-mode lowcard with two-entry buckets: parity, then timing (scattered and random keys)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_groupby_gpu.py tests/test_zz_golden_gpu.py tests/test_strkeys_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r2_pytest21.log 2>&1
tail -6 gpurun_out/r2_pytest21.log
for G in 1000 500 100; do
echo "== $G random 64-bit keys"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups $G --hashed --iters 4 2>&1 | grep "iter [13]" | cut -c1-330
done
echo "== 1000 keys, min/max set"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups 1000 --hashed --iters 3 --aggs sum,min,max,count 2>&1 | grep "iter [2]" | cut -c1-330
