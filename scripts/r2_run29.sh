#!/bin/bash
# GPU batch 29: compact partial records + hint-sized merge table: emulated ranks, single-rank communicator, chunked ingest
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_ingest_gpu.py -m gpu -q -x > gpurun_out/r2_pytest29.log 2>&1
tail -15 gpurun_out/r2_pytest29.log
