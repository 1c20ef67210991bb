"""Multi-GPU equivalence of the data-parallel training step on real GPUs (needs >= 2 devices; skipped otherwise):
tools/peer_check.py under torchrun -- fused NVLink all-reduce vs NCCL vs one GPU, graph-captured vs eager step,
in-kernel Philox increments keyed by the persistent device counter (two train() calls draw different noise and
reproduce the single-GPU stream)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_allreduce_equivalence_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "PEER_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
