"""The sharded sweeps over NCCL on a box with at least two GPUs (skipped on a single-GPU box): two ranks under torchrun,
gathered results bit-identical to the single-GPU ones (scripts/exp/dist_check.py).  The host-side sharding logic itself is
covered on CPU with gloo in tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_over_nccl_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29513", os.path.join(ROOT, "scripts", "exp", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "phsh_rel_sweep gathered == single GPU: True" in r.stdout and "cavpot_sweep gathered == single GPU: True" in r.stdout
