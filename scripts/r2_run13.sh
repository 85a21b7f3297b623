#!/bin/bash
# GPU batch 13: bucketed path, two levels + record ordering; parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_parity_large_gpu.py tests/test_groupby_gpu.py -m gpu -q -x -k "bucketed or partitioned or config2 or config5 or sharded" > gpurun_out/r2_pytest13.log 2>&1
tail -3 gpurun_out/r2_pytest13.log
for args in "--groups 262144" "--groups 1048576" "--groups 2097152" "--groups 2097152 --no-partition" "--groups 4194304" "--groups 16777216" "--groups 100000000" "--groups 100000000 --aggs sum,mean,count"; do
  echo "== $args"
  timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --iters 4 $args 2>&1 | grep "iter [13]\|Error" | cut -c1-220
done
