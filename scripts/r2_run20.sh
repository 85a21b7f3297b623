#!/bin/bash
# GPU batch 20: sort + ingest tests, merge fixes (single-rank emulation), bench with pageable e2e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_sort_gpu.py tests/test_ingest_gpu.py tests/test_multigpu_gpu.py tests/test_groupings_gpu.py -m gpu -q -x > gpurun_out/r2_pytest20.log 2>&1
tail -15 gpurun_out/r2_pytest20.log
export RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29534

unset RANK LOCAL_RANK WORLD_SIZE
(time python bench.py --no-sweep --no-extras) > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[0])
print(l['value'], l['ms_per_step'], l['e2e'])
PY
echo "== pageable e2e without the staging pipeline (driver's own bounce buffers)"
PA_H2D_STAGED=0 python bench.py --no-sweep --no-extras --no-cpu-baseline > gpurun_out/r2_bench_nostage.json 2>/dev/null
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_nostage.json').read().strip().splitlines()[0])
print(l['e2e'].get('pageable'))
PY
