#!/bin/bash
# GPU batch 48: ncu --set full of the WIDE dense k_lowcard_scan (config 2's aggregates: sum/min/max/count) at 256 and 1000 groups
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for G in 16 256 1000; do timeout 200 python scripts/prof_case.py --rows 1000000000 --groups $G --aggs sum,min,max,count --iters 3 2>&1 | grep "iter 2" | cut -c1-220; done
for G in 256 1000; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 1 -o gpurun_out/r2_lc_wide$G python scripts/prof_case.py --rows 1000000000 --groups $G --aggs sum,min,max,count --iters 1 > gpurun_out/r2_ncu_lc_wide$G.log 2>&1
ncu -i gpurun_out/r2_lc_wide$G.ncu-rep --page raw --csv > gpurun_out/r2_lc_wide${G}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lc_wide$G.ncu-rep --page source --csv > gpurun_out/r2_lc_wide${G}_src.csv 2>/dev/null
rm -f gpurun_out/r2_lc_wide$G.ncu-rep
done
ls -la gpurun_out | grep wide
