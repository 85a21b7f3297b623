#!/bin/bash
# GPU batch 50: fp32 bounds pre-check in the wide dense k_lowcard_scan: parity (group-by suite incl. the new edge-value test, large
# parity configs 1-3, golden, multi-GPU emulation), wide timings at 16 / 256 / 1000 groups, then ncu --set full of the headline
# (narrow) kernel again — lowcard.cuh changed, the traffic stamp of the bench line follows the file's hash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_groupby_gpu.py tests/test_zz_golden_gpu.py tests/test_stage2_gpu.py -m gpu -q -x > gpurun_out/r2_pytest50.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest50.log
timeout 900 python -m pytest tests/test_parity_large_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "config1 or config2_100m or fullsize or 1b or headline" 2>&1 | tail -3
for G in 16 256 1000; do timeout 200 python scripts/prof_case.py --rows 1000000000 --groups $G --aggs sum,min,max,count --iters 3 2>&1 | grep "iter 2" | cut -c1-120; done
timeout 200 python scripts/prof_case.py --rows 1000000000 --groups 1000 --iters 3 2>&1 | grep "iter 2" | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 1 -o gpurun_out/r2_lc_dense4 python scripts/prof_case.py --rows 1000000000 --groups 1000 --iters 1 > gpurun_out/r2_ncu_dense4.log 2>&1
ncu -i gpurun_out/r2_lc_dense4.ncu-rep --page raw --csv > gpurun_out/r2_lc_dense4_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 1 -o gpurun_out/r2_lc_wide1000b python scripts/prof_case.py --rows 1000000000 --groups 1000 --aggs sum,min,max,count --iters 1 > gpurun_out/r2_ncu_wide1000b.log 2>&1
ncu -i gpurun_out/r2_lc_wide1000b.ncu-rep --page raw --csv > gpurun_out/r2_lc_wide1000b_raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | grep -E "dense4|wide1000b"
