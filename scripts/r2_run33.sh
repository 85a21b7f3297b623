#!/bin/bash
# GPU batch 33: L2 prefetch in k_rp_scatter (row numbers of the tile, next tile's keys / values): launch list + sweep points
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python scripts/prof_comm1.py 1000000000 100000000 3 2>&1 | grep "^it[12]" | cut -c1-420
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_comm1_launches_b.csv python scripts/prof_comm1.py 1000000000 100000000 2 > gpurun_out/r2_comm1_ncu_b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_comm1_launches_b.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[-22:]: print(r[ki][:60], r[vi])
PY
for G in 65536 1048576 16777216 100000000; do
echo "== $G groups, sum/min/max/count"
timeout 300 python scripts/prof_bucketed.py --rows 1000000000 --groups $G --iters 3 2>&1 | grep "iter 2" | cut -c1-200
done
