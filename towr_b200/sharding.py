"""Multi-GPU sharding of a batch of problem instances.

Instances are independent, so the only thing the N > 1 path needs is a partition: one process per GPU, each
owning the contiguous range [rank*B/N, (rank+1)*B/N) with its own ``Batch``; there is no collective on the
data path.  The optional gather of the per-instance scalars (cost f64, status i32) is a plain
``torch.distributed.all_gather`` — NCCL between GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous range of instances of `rank`: [rank*total//world, (rank+1)*total//world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * total) // world, ((rank + 1) * total) // world


def shard_sizes(total, world):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_cost_status(cost, status, total, device=None):
    """All-gather the local per-instance cost / status of every rank into arrays of length `total`, in instance
    order.  `cost`: (local,) float64, `status`: (local,) int32 numpy arrays.  Requires an initialised process
    group; with world size 1 it is a copy."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.array(cost, np.float64), np.array(status, np.int32)
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total, world)
    assert len(cost) == sizes[rank] == len(status)
    pad = max(sizes)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    c = torch.zeros(pad, dtype=torch.float64, device=dev); c[:len(cost)] = torch.as_tensor(np.asarray(cost, np.float64))
    s = torch.zeros(pad, dtype=torch.int32, device=dev); s[:len(status)] = torch.as_tensor(np.asarray(status, np.int32))
    cs = [torch.empty_like(c) for _ in range(world)]
    ss = [torch.empty_like(s) for _ in range(world)]
    dist.all_gather(cs, c)
    dist.all_gather(ss, s)
    cost_all = np.concatenate([cs[r][:sizes[r]].cpu().numpy() for r in range(world)])
    status_all = np.concatenate([ss[r][:sizes[r]].cpu().numpy() for r in range(world)])
    return cost_all, status_all
