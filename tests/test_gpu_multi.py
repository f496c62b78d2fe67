"""Multi-GPU parity (needs >= 2 B200s on the box; skipped on a single-GPU box): scripts/peer_check.py under
torchrun compares the peer-memory gather fused into the kernels with a separate NCCL all-gather, bitwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_gather_equals_nccl_all_gather():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "scripts", "peer_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "peer_check ok" in out.stdout
