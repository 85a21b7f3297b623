#!/bin/bash
# GPU batch 32: merge with bucket-local home slots: parity (emulated ranks + single-rank communicator), single-GPU phases, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/r2_pytest32.log 2>&1
tail -3 gpurun_out/r2_pytest32.log
timeout 300 python scripts/prof_comm1.py 1000000000 100000000 3 2>&1 | grep "^it"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_comm1_launches.csv python scripts/prof_comm1.py 1000000000 100000000 2 > gpurun_out/r2_comm1_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_comm1_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
out=[(r[ki][:60], r[vi], r[ui]) for r in rows[1:]]
# last step only: from the last k_rp_* / k_bucket_agg backwards
names=[o[0] for o in out]
last=max(i for i,nm in enumerate(names) if nm.startswith('k_bucket_agg'))
start=max(i for i,nm in enumerate(names[:last]) if nm.startswith('k_rp_scatter')) - 3
for o in out[max(start,0):]: print(o)
PY
