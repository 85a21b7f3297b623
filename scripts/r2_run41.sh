#!/bin/bash
# GPU batch 41: resample scan with the one-vote steady-state path: parity, then config 4 timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_resample_gpu.py tests/test_zz_golden_gpu.py tests/test_facade_gpu.py tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/r2_pytest41.log 2>&1
tail -3 gpurun_out/r2_pytest41.log
timeout 600 python -m pytest tests/test_parity_large_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "config4 or resample" 2>&1 | tail -3
timeout 300 python scripts/prof_resample.py --rows 1000000000 --iters 3 2>&1 | grep "iter [12]" | cut -c1-200
timeout 300 python scripts/prof_resample.py --rows 1000000000 --iters 3 --aggs sum,mean,count 2>&1 | grep "iter 2" | cut -c1-200
