#!/bin/bash
# GPU batch 5 (2 GPUs): utf8-key tests, real NCCL parity, N=2 bench with config 5 through pa_comm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_strkeys_gpu.py tests/test_facade_gpu.py tests/test_zz_golden_gpu.py -m gpu -q -x > gpurun_out/r2_pytest5a.log 2>&1
tail -3 gpurun_out/r2_pytest5a.log
timeout 900 python -m pytest tests/test_nccl_multigpu.py -m gpu -q -x -s > gpurun_out/r2_pytest5b.log 2>&1
tail -15 gpurun_out/r2_pytest5b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench5_n2.json 2> gpurun_out/r2_bench5_n2.err
tail -c 1500 gpurun_out/r2_bench5_n2.json
