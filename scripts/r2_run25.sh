#!/bin/bash
# GPU batch 25 (2 GPUs): headline at N=2 with 0 / 1 / 2 / 4 SMs left free by the scan
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for R in 0 1 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$R bench.py --gpus 2 --steps 20 --warmup 4 --no-e2e --config5-groups 0 --sm-reserve $R > gpurun_out/r2_bench_n2_res$R.json 2> gpurun_out/r2_bench_n2_res$R.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_n2_res$R.json').read().strip().splitlines()[-1])
print($R, l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['config']['groups_found_global'])
PY
done
