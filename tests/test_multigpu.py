"""GPU, N >= 2: the row-sharded half-iteration with either exchange of the solved rows, through torchrun (skipped on 1 GPU)."""
import os
import socket
import subprocess
import sys

import pytest

from rsparse_b200 import _lib as L

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_sharded_half_iteration_vs_oracle(exchange):
    """exchange = p2p: peer-memory pushes (a silent NCCL fallback is an error); nccl: grouped broadcasts."""
    n = L.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    env = dict(os.environ, B200ALS_EXCHANGE=exchange)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MULTIGPU_OK world=2" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
