"""Real multi-process NCCL parity: torchrun with 2 ranks (one per GPU); the gathered result of every
transport — the C-ABI communicator (pa_groupby_sharded_aggregate), the counted and the padded exchange of
distributed.py — is compared with the oracle on the whole data set (tests/nccl_worker.py).  Needs 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_nccl_multigpu.py -m gpu`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2])
def test_nccl_ranks_match_oracle(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0 and "NCCL PARITY OK" in r.stdout
