"""Multi-GPU host logic: independent streams (images) are sharded over one process per GPU, exactly like the
reference shards its dataloader under `accelerate launch` (src/inference.py:82) -- there is NO collective on the
decode path.  torch.distributed is used only to agree on the timing (barrier + max over ranks) and to gather
per-rank counts after the timed region.
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend=None):
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def rank_world():
    """(rank, world) of the initialised process group, (0, 1) for a single process."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items, rank, world):
    """Static round-robin partition: stream i -> rank i mod world."""
    return [it for i, it in enumerate(items) if i % world == rank]


def barrier():
    if dist.is_initialized():
        dist.barrier()


def reduce_max(value, device=None):
    """max over ranks of a python float (timing must be the slowest rank's)."""
    if not dist.is_initialized():
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, device=None):
    if not dist.is_initialized():
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
