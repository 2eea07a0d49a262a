"""Multi-GPU use of the query path: queries are independent, so they shard with no data-path
collective.  Tables are replicated -- broadcast once at set-up (NCCL over NVLink/NVSwitch on
GPUs, gloo in the CPU tests) -- and every rank evaluates a contiguous slice of the queries.
Timing across ranks is the maximum over ranks (`max_over_ranks`).

In JAX terms this is shard_map over a 1-D mesh with queries P("q") and tables P().
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of n items for `rank` of `world`; slices start at
    even offsets so that the pair-per-lane kernel keeps its 16-byte aligned loads."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request")
    per = -(-n // world)
    per += per & 1
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def broadcast_tables(tensors, src: int = 0) -> None:
    """One-time replication of the tables (no-op without an initialised process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def max_over_ranks(value: float, device=None) -> float:
    """Max-reduce a host scalar over the process group (how multi-GPU timings are reported)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
