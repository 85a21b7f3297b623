#!/bin/bash
# GPU batch 38: ncu --set full of the sharded step's own kernels (single-rank communicator, 1 B rows, 100 M groups)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 ncu --set full --clock-control none -k regex:"k_bkrec_count|k_bkrec_scatter|k_merge_insert|k_merge_compact|k_merge_fold" -s 5 -c 5 -o gpurun_out/r2_sharded -f python scripts/prof_comm1.py 1000000000 100000000 2 > gpurun_out/r2_sharded.log 2>&1
ncu -i gpurun_out/r2_sharded.ncu-rep --page raw --csv > gpurun_out/r2_sharded_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_sharded_raw.csv > gpurun_out/r2_sharded.md
rm -f gpurun_out/r2_sharded.ncu-rep
grep -E "^## |time_duration|dram__bytes|lts__t_sectors_op_atom" gpurun_out/r2_sharded.md | cut -c1-120
