"""Host-side multi-rank logic on CPU with the gloo backend (world_size 2): balanced disjoint
query shards, one-time table broadcast, max-over-ranks timing.  No GPU needed."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interpax_b200.sharding import broadcast_tables, max_over_ranks, shard_bounds


def test_shard_bounds_cover_without_overlap():
    for n in (0, 1, 2, 7, 64, 1000, 10**8 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (l0, h0), (l1, h1) in zip(spans, spans[1:]):
                assert h0 == l1 and l0 <= h0
            assert all(lo % 2 == 0 or lo == n for lo, _ in spans)
            sizes = [h - l for l, h in spans]
            assert max(sizes) <= -(-n // world) + 1  # balanced: at most one over the even split
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # tables: rank 0 owns the data, everyone else starts from garbage
        g = torch.Generator().manual_seed(7)
        ref = torch.randn(16, 16, generator=g, dtype=torch.float64)
        tab = ref.clone() if rank == 0 else torch.full_like(ref, float("nan"))
        broadcast_tables([tab])
        assert torch.equal(tab, ref)
        # queries: each rank takes its slice; together they cover everything exactly once
        n = 1001
        lo, hi = shard_bounds(n, world, rank)
        mine = torch.zeros(n, dtype=torch.int64)
        mine[lo:hi] = 1
        dist.all_reduce(mine)
        assert torch.equal(mine, torch.ones(n, dtype=torch.int64))
        # timing: the slowest rank defines the step time
        t = max_over_ranks(1.0 + rank)
        assert t == float(world)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    spans = [np.load(tmp_path / f"ok{r}.npy") for r in range(world)]
    assert spans[0][1] == spans[1][0]
