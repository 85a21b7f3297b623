#!/bin/bash
# GPU batch 42: fresh rebuild of the tree -> whole GPU suite, then the default bench line (refreshes profiles/r2_bench_line.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1100 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2_pytest42.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest42.log
timeout 400 python bench.py > gpurun_out/r2_bench42.json 2> gpurun_out/r2_bench42.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench42.json
