"""Multi-GPU plumbing of the hot path: row sharding + the single mean-NLL collective.

The reference has no distributed code; the path shards trivially by rows
(every op is row-wise, `zuko/nn.py:217-218`, `zuko/transforms.py:554-567`), so the only
exchange is ONE all-reduce of ``{sum log p, count}`` — two doubles — per step.
One process per GPU (``torchrun``), ``torch.distributed`` backend ``nccl`` (``gloo`` in the
CPU tests).
"""

from __future__ import annotations

__all__ = ["mean_nll", "shard_rows"]

import torch
import torch.distributed as dist
from torch import Tensor


def shard_rows(n_rows: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced row range ``[start, stop)`` of rank ``rank`` (the first
    ``n_rows % world_size`` ranks get one extra row)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_rows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def mean_nll(sum_log_prob: Tensor, count: int, group: dist.ProcessGroup | None = None) -> Tensor:
    """Global mean negative log-likelihood from each rank's ``sum(log_prob)`` (a double
    tensor with one element, as returned by ``NormalizingFlow.log_prob_and_sum``) and row
    count: one ``all_reduce(sum)`` over a 2-element double buffer."""
    buf = torch.empty(2, dtype=torch.float64, device=sum_log_prob.device)
    buf[0] = sum_log_prob.reshape(-1)[0].to(torch.float64)
    buf[1] = float(count)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return -(buf[0] / buf[1])
