#!/bin/bash
# GPU batch 17 (2 GPUs): N=2 bench line with two-lane step pipelining (headline only, no config 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for i in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 4 --no-e2e --config5-groups 0 > gpurun_out/r2_bench_n2_lanes$i.json 2> gpurun_out/r2_bench_n2_lanes$i.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_n2_lanes$i.json').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['config']['groups_found_global'], l['gpu_launches'])
PY
tail -3 gpurun_out/r2_bench_n2_lanes$i.err
done
