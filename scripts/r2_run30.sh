#!/bin/bash
# GPU batch 29-30: compact partial records, hint-sized merge table, unordered bucket-record export: emulated ranks + single-rank communicator
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/r2_pytest30.log 2>&1
tail -15 gpurun_out/r2_pytest30.log
