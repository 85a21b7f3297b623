#!/bin/bash
# GPU batch 11: ncu --set full of the level-1 scatter at two fan-outs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for B in 5 8; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rp_scatter -c 1 -o gpurun_out/r2_scatter_b$B -f python scripts/prof_bucketed.py --rows 1000000000 --iters 1 --groups 65536 --bits $B > gpurun_out/r2_scatter_b$B.log 2>&1
ncu -i gpurun_out/r2_scatter_b$B.ncu-rep --page raw --csv > gpurun_out/r2_scatter_b${B}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_scatter_b${B}_raw.csv > gpurun_out/r2_scatter_b${B}.md
done
