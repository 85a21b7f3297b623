#!/bin/bash
# GPU batch 35 (N GPUs): the N-GPU bench line with config 5 (no e2e / cpu legs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n${N}_c.json 2> gpurun_out/r2_bench_n${N}_c.err
tail -c 1500 gpurun_out/r2_bench_n${N}_c.json; tail -3 gpurun_out/r2_bench_n${N}_c.err
