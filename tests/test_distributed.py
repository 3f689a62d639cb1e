"""CPU, world_size = 2, gloo: the host-side logic of the multi-GPU path (row sharding, id broadcast,
max-over-ranks timing, re-assembly of row blocks).  The data path on each rank is the CPU oracle here;
on GPUs the same plumbing drives libb200als.so over NCCL (bench.py --gpus N)."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch, torch.distributed as dist
import oracle, wrmf_cases as wc
from rsparse_b200 import parallel
rank, world = parallel.init_process_group("gloo")
assert world == 2
# 1. the 128-byte id travels intact
payload = bytes(range(128)) if rank == 0 else b"\0" * 128
assert parallel.broadcast_bytes(payload, 0) == bytes(range(128))
# 2. max-over-ranks
assert parallel.max_over_ranks(10.0 + rank) == 11.0
# 3. row-sharded half-iteration == single-process half-iteration
c = wc.half_iteration_cases()["synth_ragged_implicit_cg_k128"]
n = c["Y0"].shape[0]
b, e = parallel.shard_range(n, rank, world)
ptr = (c["ptr"][b:e + 1] - c["ptr"][b]).astype(np.int32)
sl = slice(c["ptr"][b], c["ptr"][e])
X = c["X"]; G = oracle.gram(X, c["lam"], 1)
Yloc = c["Y0"][b:e].copy()
num = oracle.als_implicit(ptr, c["idx"][sl], c["val"][sl], X, Yloc, G, 0.0, wc.CG, 3, 1) * max(1, int(ptr[-1]))
# exchange: every rank contributes its block (unequal sizes -> one broadcast per owner, as the engine does)
full = np.zeros_like(c["Y0"])
for r in range(world):
    rb, re = parallel.shard_range(n, r, world)
    t = torch.from_numpy(Yloc.copy() if r == rank else np.zeros((re - rb, X.shape[1]), np.float32))
    dist.broadcast(t, src=r)
    full[rb:re] = t.numpy()
ref = c["Y0"].copy()
oracle.als_implicit(c["ptr"], c["idx"], c["val"], X, ref, G, 0.0, wc.CG, 3, 1)
assert np.array_equal(full, ref)
parallel.barrier()
print("rank", rank, "ok")
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_row_sharding(tmp_path):
    port = _free_port()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def test_shard_helpers():
    from rsparse_b200 import parallel
    for n, w in ((10, 3), (10_000_000, 8), (7, 8), (0, 2)):
        cuts = [parallel.shard_range(n, r, w) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in cuts]
        assert max(sizes) - min(sizes) <= 1
    ptr = np.concatenate([[0], np.cumsum(np.r_[np.full(100, 1), np.full(10, 1000)])])
    cuts = parallel.shard_by_nnz(ptr, 4)
    assert cuts[0] == 0 and cuts[-1] == 110 and np.all(np.diff(cuts) >= 0)
    per = [ptr[cuts[i + 1]] - ptr[cuts[i]] for i in range(4)]
    assert max(per) <= ptr[-1] / 4 + 1000
