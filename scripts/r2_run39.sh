#!/bin/bash
# GPU batch 39: 8192-row scatter tiles for fan-outs >= 128 ways: parity of the bucketed path, then A/B at 16 M / 100 M groups
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_parity_large_gpu.py -m gpu -q -x -k "bucketed or partition" > gpurun_out/r2_pytest39.log 2>&1
tail -3 gpurun_out/r2_pytest39.log
for G in 16777216 100000000; do
for T in 99 8 7; do
echo "== $G groups, sum/min/max/count, big tiles from 2^$T ways"
PA_RP_BIG_TILE_LOG=$T timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --groups $G --iters 3 2>&1 | grep "iter 2" | cut -c1-120
done
done
echo "== 100 M groups sum/mean/count (narrow), big from 99 / 7"
PA_RP_BIG_TILE_LOG=99 timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --groups 100000000 --aggs sum,mean,count --iters 3 2>&1 | grep "iter 2" | cut -c1-120
PA_RP_BIG_TILE_LOG=7 timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --groups 100000000 --aggs sum,mean,count --iters 3 2>&1 | grep "iter 2" | cut -c1-120
