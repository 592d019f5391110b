"""Host-side multi-rank logic on CPU: world_size-2 gloo group (no GPU needed)."""

import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from lumol_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    def broadcast(payload):
        tensor = torch.zeros(128, dtype=torch.uint8)
        if payload is not None:
            tensor = torch.tensor(list(payload), dtype=torch.uint8)
        dist.broadcast(tensor, 0)
        return bytes(tensor.tolist())

    unique = parallel.exchange_unique_id(rank, lambda: bytes(range(128)), broadcast)
    slowest = parallel.max_over_ranks(10.0 + rank, world)
    # strong-scaling partition: blocks are contiguous, equal-sized (all-gather friendly) and cover every atom
    n = 1000003
    lo, hi = parallel.owned_range(n, rank, world)
    covered = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(covered)
    results[rank] = (unique == bytes(range(128)), slowest, int(covered.item()) == n, lo, hi)
    dist.destroy_process_group()


def test_two_rank_bootstrap_and_partition():
    world = 2
    port = _free_port()
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    for rank in range(world):
        same_id, slowest, covered, lo, hi = results[rank]
        assert same_id and covered
        assert slowest == 11.0
    assert results[0][4] == results[1][3]  # contiguous


def test_owned_range_matches_allgather_blocks():
    for n in (0, 1, 7, 1000, 1048576, 1048577):
        for world in (1, 2, 3, 4, 8):
            chunk = (n + world - 1) // world
            total = 0
            for rank in range(world):
                lo, hi = parallel.owned_range(n, rank, world)
                assert 0 <= lo <= hi <= n and hi - lo <= chunk
                assert lo == min(n, rank * chunk)
                total += hi - lo
            assert total == n
