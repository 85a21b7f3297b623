#!/bin/bash
# GPU batch 36 (N GPUs): bench line WITH the e2e leg (NUMA-bound pinned host columns), no config 5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m 2>/dev/null | head -14
for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done
lscpu | grep -i "numa\|socket\|^CPU(s)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --config5-groups 0 > gpurun_out/r2_bench_n${N}_e2e.json 2> gpurun_out/r2_bench_n${N}_e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n${N}_e2e.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], json.dumps(d['e2e']))
PY
tail -3 gpurun_out/r2_bench_n${N}_e2e.err
