#!/bin/bash
# GPU batch 9: bucketed path after the atomic-free partition rewrite: parity, then per-kernel times (ncu launch list)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_parity_large_gpu.py tests/test_groupby_gpu.py -m gpu -q -x -k "bucketed or partitioned or config2" > gpurun_out/r2_pytest9.log 2>&1
tail -5 gpurun_out/r2_pytest9.log
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
half=len(body)//2
tot=0
for r in body[-(len(body)-2)//2:]:
    if len(r)>vi:
        print('  ', r[ki][:50], round(float(r[vi])/1e6,3),'ms'); tot+=float(r[vi])/1e6
print('   total', round(tot,2))
PY
  grep "iter 1" gpurun_out/r2_bk_$tag.log | cut -c1-250
}
run 64K --groups 65536
run 1M --groups 1048576
run 1M_2lvl --groups 1048576 --bits 0x0505
run 16M --groups 16777216
run 100M --groups 100000000
