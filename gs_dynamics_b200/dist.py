"""Multi-GPU plumbing: one process per GPU (torchrun), episodes / rollouts shard across ranks with NO data-path collective
(tracking episodes and GNN rollouts share nothing, train_gs.py:14-18).  torch.distributed (NCCL on GPUs, gloo in
the CPU tests) is used only for the barrier, the max-over-ranks timing reduction and scalar statistics."""
import os

import torch
import torch.distributed as dist


def env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None, device=None):
    rank, local_rank, world = env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def shard(n_units, rank, world):
    """Units (episodes, rollouts) of this rank: contiguous blocks, sizes differing by at most one."""
    base, rem = divmod(n_units, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def barrier():
    if dist.is_initialized():
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def reduce_max(values, device=None):
    """Max over ranks of a list of python floats (device timings)."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def reduce_sum(values, device=None):
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def finalize():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
