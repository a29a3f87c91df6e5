"""Multi-GPU parity on real GPUs (pytest -m gpu; skipped with fewer than two devices): torchrun with 2 ranks over
NCCL — the row-sharded VQ path equals the one-GPU result (indices and z_q per shard bit-identical, histogram exactly
equal, loss / perplexity within 1e-6), for the FP32 and the tcgen05 kernel.  The worker is scripts/dist_parity.py."""
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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dist_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "DIST_PARITY PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
