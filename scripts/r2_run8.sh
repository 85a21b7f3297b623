#!/bin/bash
# GPU batch 8: per-kernel times of the bucketed path (ncu launch list) at 1 B rows
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for G in 65536 1048576 100000000; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_bk_launch_$G.csv python scripts/prof_bucketed.py --rows 1000000000 --groups $G --iters 2 > gpurun_out/r2_bk_$G.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2_bk_launch_$G.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
print('G=$G')
for r in rows[hdr+1:]:
    if len(r)>vi: print('  ', r[ki][:60], r[vi], r[ui])
PY
done
