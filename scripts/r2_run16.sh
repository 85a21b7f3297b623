#!/bin/bash
# GPU batch 16 (2 GPUs): real NCCL parity test, then the N=2 bench line (headline + config 5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi -L
timeout 900 python -m pytest tests/test_nccl_multigpu.py tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/r2_pytest16.log 2>&1
tail -5 gpurun_out/r2_pytest16.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 2500 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
