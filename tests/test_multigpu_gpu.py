"""Two B200s, NCCL: the view-sharded training step's all-reduced gradients equal the mean of the two shards run serially
(1e-5 relative, or the serial computation's own run-to-run noise from float atomics if that is larger).  Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_rank_sharded_step_matches_serial_shards():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "two_rank_step.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", script], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], "\n".join(ln for ln in r.stderr.splitlines() if "max err" in ln or "Error" in ln))
    assert r.returncode == 0 and "TWO_RANK_OK" in r.stdout
