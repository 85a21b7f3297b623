#!/bin/bash
# GPU batch 15: calendar rules (parity + facade), smem-front kernel at 1000 scattered keys
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_resample_gpu.py tests/test_facade_gpu.py -m gpu -q -x > gpurun_out/r2_pytest15.log 2>&1
tail -15 gpurun_out/r2_pytest15.log
for p in auto global; do
echo "== 1000 scattered keys path=$p"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups 1000 --hashed --iters 4 --path $p 2>&1 | grep "iter [13]" | cut -c1-330
done
echo "== 1000 dense keys path=global"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups 1000 --iters 4 --path global 2>&1 | grep "iter [13]" | cut -c1-330
