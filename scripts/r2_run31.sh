#!/bin/bash
# GPU batch 31 (2 GPUs): real NCCL parity (incl. compact records), then the N=2 bench line with config 5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi -L
timeout 600 python -m pytest tests/test_nccl_multigpu.py -m gpu -q -x > gpurun_out/r2_pytest31.log 2>&1
tail -5 gpurun_out/r2_pytest31.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err
tail -c 1800 gpurun_out/r2_bench_n2_b.json; tail -3 gpurun_out/r2_bench_n2_b.err
