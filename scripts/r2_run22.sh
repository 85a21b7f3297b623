#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
M=hash2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -s 1 -c 1 -o gpurun_out/r2_lc_$M -f python scripts/prof_case.py --rows 1000000000 --groups 1000 --iters 2 --hashed > gpurun_out/r2_lc_$M.log 2>&1
ncu -i gpurun_out/r2_lc_$M.ncu-rep --page raw --csv > gpurun_out/r2_lc_${M}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lc_$M.ncu-rep --page source --csv > gpurun_out/r2_lc_${M}_src.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_lc_${M}_raw.csv > gpurun_out/r2_lc_${M}.md
rm -f gpurun_out/r2_lc_$M.ncu-rep
