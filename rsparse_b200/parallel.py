"""Host-side multi-GPU plumbing: one process per GPU (torchrun), rows sharded by contiguous
blocks, NCCL communicator owned by libb200als.so.  torch.distributed is used
only to hand the 128-byte NCCL id around and for barriers / max-over-ranks of timings."""
import os

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [begin, end) of n rows owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_by_nnz(ptr, world):
    """Contiguous row blocks balanced by the nnz prefix sum (for skewed matrices): returns world+1 cut points."""
    ptr = np.asarray(ptr, dtype=np.int64)
    n = len(ptr) - 1
    targets = ptr[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(ptr, targets, side="left")
    cuts = np.clip(cuts, 0, n)
    return np.concatenate([[0], np.maximum.accumulate(cuts), [n]]).astype(np.int64)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend="gloo"):
    """Join the torchrun rendezvous (MASTER_ADDR/MASTER_PORT from the environment)."""
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` to every rank over torch.distributed."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    n = 128
    t = torch.zeros(n, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    dist.broadcast(t, src=src)
    return bytes(t.numpy().tobytes())


def init_engine_comm():
    """Create the engine's NCCL communicator across the torchrun ranks (call after set_device)."""
    import ctypes as C

    from . import _lib as L
    rank, world, _ = env_rank_world()
    if world == 1:
        return rank, world       # single GPU: neither torch nor NCCL is touched
    rank, world = init_process_group("gloo")
    buf = (C.c_char * 128)()
    if rank == 0:
        L.check(L.lib().b200als_comm_unique_id(buf))
    uid = broadcast_bytes(bytes(buf.raw), 0)
    buf2 = (C.c_char * 128).from_buffer_copy(uid)
    L.check(L.lib().b200als_comm_init(buf2, rank, world))
    return rank, world


def max_over_ranks(x):
    if env_rank_world()[1] == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if env_rank_world()[1] == 1:
        return
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def bind_to_gpu_numa(local_rank):
    """Pin this process (and, by first touch, the pinned host buffers it allocates afterwards) to the CPUs of the NUMA node
    the GPU hangs off: with one process per GPU the host<->device copies of the stateless calls then read / write local
    DRAM instead of crossing the socket interconnect.  Best effort: returns a description, never raises."""
    import subprocess
    try:
        bdf = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bdf:
            return "numa: no pci id"
        if len(bdf.split(":")[0]) == 8:            # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return "numa: single node (gpu %s)" % bdf
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return "numa: node %d has no allowed cpus" % node
        os.sched_setaffinity(0, allowed)
        return "numa: gpu %s -> node %d (%d cpus)" % (bdf, node, len(allowed))
    except Exception as e:  # noqa: BLE001
        return "numa: not bound (%s)" % type(e).__name__
