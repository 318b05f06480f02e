"""Multi-GPU plumbing of the hot path: row sharding + the single mean-NLL collective.

The reference has no distributed code; the path shards trivially by rows
(every op is row-wise, `zuko/nn.py:217-218`, `zuko/transforms.py:554-567`), so the only
exchange of the forward path is ONE all-reduce of ``{sum log p, count}`` — two doubles — per step;
data-parallel TRAINING adds the usual gradient all-reduce (``all_reduce_gradients``, or wrap the flow
in ``torch.nn.parallel.DistributedDataParallel``: the engine's gradients arrive through ordinary
autograd ``AccumulateGrad`` nodes, so DDP's bucket hooks work unchanged).
One process per GPU (``torchrun``), ``torch.distributed`` backend ``nccl`` (``gloo`` in the
CPU tests).
"""

from __future__ import annotations

__all__ = ["all_reduce_gradients", "mean_nll", "shard_rows"]

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


def all_reduce_gradients(module: torch.nn.Module, weights: tuple[int, int] | None = None,
                         group: dist.ProcessGroup | None = None) -> None:  # fmt: skip
    """Data-parallel gradient reduction for a manual training loop: ONE ``all_reduce(sum)`` over a flat
    buffer of every ``.grad`` of ``module`` (parameters without a gradient contribute zeros), written
    back in place.  With ``weights = (local_rows, global_rows)`` each rank's gradient of its LOCAL
    mean loss is re-weighted by ``local_rows / global_rows`` first, so that the result is the gradient
    of the global mean NLL also for ragged shards (``shard_rows``); without it gradients are
    averaged over the ranks."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params:
        return
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    scale = (weights[0] / weights[1]) if weights is not None else 1.0 / world
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in params]) * scale
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in params:
        n = p.numel()
        p.grad = flat[off : off + n].reshape(p.shape).to(p.dtype).clone()
        off += n
