#!/bin/bash
# GPU batch 40: nullable value columns on the bucketed path: parity (new test, config 3 at 100 M rows, nullable cases of the main suite), config 3 timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 1200 python -m pytest tests/test_parity_large_gpu.py tests/test_groupby_gpu.py tests/test_stage2_gpu.py -m gpu -q -x > gpurun_out/r2_pytest40.log 2>&1
tail -5 gpurun_out/r2_pytest40.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['extras']['multikey_nullable']))"
