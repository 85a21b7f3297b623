#!/bin/bash
# GPU batch 6: bucketed path (tests vs the oracle, then the whole suite), bench with sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_parity_large_gpu.py -m gpu -q -x -k "bucketed or 16m" > gpurun_out/r2_pytest6a.log 2>&1
tail -5 gpurun_out/r2_pytest6a.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
for s in d.get('sweep',[]): print(s['groups'], s['path'], round(s['total_ms'],2), round(s['scan_ms'],2), round(s['frac'],3))
PY
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6b.log 2>&1
tail -5 gpurun_out/r2_pytest6b.log
