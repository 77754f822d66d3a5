"""CPU tests of the N > 1 path: the instance partition and the optional gather of per-instance cost / status,
with two gloo processes (the GPU run uses the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest

from towr_b200.sharding import gather_cost_status, shard_range, shard_sizes


def test_partition_is_contiguous_and_complete():
    for total in (1, 7, 64, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    idx = np.arange(lo, hi)
    cost = idx * 0.5 + 1.0                       # stands for the per-instance cost of this rank's shard
    status = (idx % 5 == 0).astype(np.int32)
    cost_all, status_all = gather_cost_status(cost, status, total)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), cost=cost_all, status=status_all)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_cost_status_two_gloo_ranks(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    total = 67                                    # uneven shards: 33 + 34
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    idx = np.arange(total)
    for rank in range(2):
        z = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(z["cost"], idx * 0.5 + 1.0)
        assert np.array_equal(z["status"], (idx % 5 == 0).astype(np.int32))
