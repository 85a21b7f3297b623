#!/bin/bash
# GPU batch 12: bucketed path: ranged mode, cached plan; parity; comparison with the plain L2-atomic path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_parity_large_gpu.py tests/test_groupby_gpu.py -m gpu -q -x > gpurun_out/r2_pytest12.log 2>&1
tail -5 gpurun_out/r2_pytest12.log
for args in "--groups 65536" "--groups 65536 --no-partition" "--groups 262144" "--groups 262144 --no-partition" "--groups 1048576" "--groups 1048576 --no-partition" "--groups 4194304" "--groups 4194304 --no-partition" "--groups 16777216" "--groups 65536 --scattered" "--groups 1048576 --aggs sum,mean,count" "--groups 1048576 --aggs sum,mean,count --no-partition"; do
  echo "== $args"
  timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --iters 4 $args 2>&1 | grep "iter [13]" | cut -c1-220
done
