#!/bin/bash
# GPU batch 10: fan-out sweep of the single-level bucketed path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
run() {  # tag, args...
  tag=$1; shift
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_bk_launch_$tag.csv python scripts/prof_bucketed.py --rows 1000000000 --iters 2 "$@" > gpurun_out/r2_bk_$tag.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2_bk_launch_$tag.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
print('== $tag $*')
body=rows[hdr+1:]
tot=0
for r in body[-(len(body)-2)//2:]:
    if len(r)>vi:
        ms=float(r[vi])/1e6; tot+=ms
        if ms>0.3: print('  ', r[ki][:50], round(ms,3),'ms')
print('   total', round(tot,2))
PY
}
run 64K_b7 --groups 65536 --bits 7
run 64K_b8 --groups 65536 --bits 8
run 64K_b9 --groups 65536 --bits 9
run 1M_b9 --groups 1048576 --bits 9
run 1M_narrow_b8 --groups 1048576 --bits 8 --aggs sum,mean,count
run 1M_narrow_b9 --groups 1048576 --bits 9 --aggs sum,mean,count
