#!/bin/bash
# GPU batch 45 (2 GPUs): the real 2-rank NCCL parity test, then the bench exactly as the driver launches it at N = 2 (all legs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_nccl_multigpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_d.json 2> gpurun_out/r2_bench_n${N}_d.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench_n${N}_d.json; tail -3 gpurun_out/r2_bench_n${N}_d.err
