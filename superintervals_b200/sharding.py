"""Multi-GPU sharding of the batch query path: one process per GPU.

Queries are independent and the index is read-only after build(), so the path shards
with NO data-path collective:

  mode A (replicated index)   every rank builds/holds the whole index and answers a
                              contiguous slice of the query batch;
  mode B (contig-partitioned) one index per contig (the reference's own convention,
                              examples/bed-intersect-si.rs:100-123), contigs assigned to
                              ranks by longest-processing-time on (N_c + Q_c).

The only exchange is the CSR bookkeeping: each rank's total hit count is all-gathered
so every rank knows the base offset of its value segment in the global CSR (a rank's
segment is contiguous because its query range is). Backend-agnostic: NCCL over
NVLink on the B200 box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .workloads import lpt_assign, shard_range

__all__ = ["shard_range", "lpt_assign", "csr_shard_bases", "gather_counts", "assign_contigs"]


def csr_shard_bases(local_total_hits: int, device=None, group=None):
    """All-gather per-rank hit totals. Returns (bases[world], totals[world]) as Python ints:
    rank r's values occupy [bases[r], bases[r] + totals[r]) of the global CSR."""
    if not (dist.is_available() and dist.is_initialized()):
        return [0], [int(local_total_hits)]
    world = dist.get_world_size(group)
    mine = torch.tensor([int(local_total_hits)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    totals = [int(t.item()) for t in out]
    bases = [0] * world
    for r in range(1, world):
        bases[r] = bases[r - 1] + totals[r - 1]
    return bases, totals


def gather_counts(local_counts: torch.Tensor, n_total: int, group=None):
    """Mode A: assemble the full per-query count vector on every rank from contiguous
    slices (shard_range order). Slices may differ by one element, so pad to the max."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_counts
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    buf[: local_counts.numel()] = local_counts
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])


def assign_contigs(n_intervals, n_queries, world):
    """Mode B: owner rank of every contig, LPT on (N_c + Q_c)."""
    cost = np.asarray(n_intervals, np.float64) + np.asarray(n_queries, np.float64)
    return lpt_assign(cost, world)
