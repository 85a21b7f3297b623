#!/bin/bash
# GPU batch 18: kernel breakdown of the counted export + merge at 100 M groups (one rank emulating the owner of everything)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
export RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29533
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_merge_launch.csv python scripts/prof_dist.py 1000000000 100000000 > gpurun_out/r2_merge.log 2>&1
tail -4 gpurun_out/r2_merge.log | cut -c1-300
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2_merge_launch.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
body=[r for r in rows[hdr+1:] if len(r)>vi]
# last iteration: from the last k_rp_scatter-or-hist backwards... print the last 45 launches
for r in body[-45:]:
    ms=float(r[vi])/1e6
    print('  ', r[ki][:70], round(ms,3),'ms')
PY
