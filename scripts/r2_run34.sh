#!/bin/bash
# GPU batch 34: ncu --set full of the current hash-mode and dense-mode k_lowcard_scan (1 B rows, 1000 keys, sum/mean/count)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups 1000 --hashed --iters 3 2>&1 | grep "iter 2" | cut -c1-160
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -s 1 -c 1 -o gpurun_out/r2_lc_hash3 python scripts/prof_case.py --rows 1000000000 --groups 1000 --hashed --iters 1 > gpurun_out/r2_ncu_hash3.log 2>&1
ncu -i gpurun_out/r2_lc_hash3.ncu-rep --page raw --csv > gpurun_out/r2_lc_hash3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lc_hash3.ncu-rep --page source --csv > gpurun_out/r2_lc_hash3_src.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 1 -o gpurun_out/r2_lc_dense3 python scripts/prof_case.py --rows 1000000000 --groups 1000 --iters 1 > gpurun_out/r2_ncu_dense3.log 2>&1
ncu -i gpurun_out/r2_lc_dense3.ncu-rep --page raw --csv > gpurun_out/r2_lc_dense3_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lc_dense3.ncu-rep --page source --csv > gpurun_out/r2_lc_dense3_src.csv 2>/dev/null
rm -f gpurun_out/r2_lc_hash3.ncu-rep gpurun_out/r2_lc_dense3.ncu-rep
ls -la gpurun_out/ | grep hash3
