"""Sharded state over NCCL on >= 2 GPUs of one box vs the CPU oracle (skipped on a 1-GPU box;
the host logic is covered on CPU by tests/test_sharded_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_nccl_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    ng = 4 if torch.cuda.device_count() >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ng),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "check_sharded.py"),
           "--qubits", "18"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "SHARDED_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
