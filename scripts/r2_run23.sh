#!/bin/bash
# GPU batch 23: facade tests (apply_chunk), ncu of the scatter-shaped take + counting sort, DSMEM micro-benchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_facade_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 600 ncu --set full --clock-control none -k regex:"k_take_scatter|k_cs_scatter|k_rowid_scan" -c 3 -o gpurun_out/r2_groupings -f python scripts/prof_groupings.py --rows 200000000 --groups 1000 --iters 1 > gpurun_out/r2_groupings.log 2>&1
ncu -i gpurun_out/r2_groupings.ncu-rep --page raw --csv > gpurun_out/r2_groupings_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_groupings_raw.csv > gpurun_out/r2_groupings.md
rm -f gpurun_out/r2_groupings.ncu-rep
grep -E "^## |time_duration|dram__bytes" gpurun_out/r2_groupings.md | cut -c1-120
timeout 120 ./scripts/ubench/_bin/dsmem_atomics 2>&1 | tail -12
