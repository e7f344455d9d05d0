"""Data parallelism on GPUs: FlatDataParallel + FlatSGD (bf16 gradient payload, reduce-scattered AVT-h gradients, sharded
update, all-gathered bf16 weights) must leave the same weights as stock autograd + torch SGD on the global batch."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_dp.py")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "check_dp" in r.stdout


@pytest.mark.gpu
def test_two_ranks_on_one_gpu_equal_global_batch_step():
    """Runs on the single-GPU box too: two processes share cuda:0 and exchange over gloo."""
    _run(["--one-gpu"])


@pytest.mark.gpu
def test_two_rank_nccl_step_equals_global_batch_step():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run([])
