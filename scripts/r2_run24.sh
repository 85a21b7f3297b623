#!/bin/bash
# GPU batch 24 (8 GPUs): the N=8 bench line (headline with the tail stream + config 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err | cut -c1-300
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print(l['n_gpus'], l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['config']['groups_found_global'], l['e2e'])
print(json.dumps(l.get('extras'))[:900])
PY
