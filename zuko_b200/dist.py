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

__all__ = ["NllRing", "all_reduce_gradients", "mean_nll", "shard_rows"]

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


class NllRing:
    """The mean-NLL collective taken off the critical path.

    ``mean_nll`` issues a blocking 16-byte all-reduce (plus a few eager fills) on the compute
    stream every step, so every step costs the max over ranks plus the collective's latency.
    The ring keeps the per-step terms ``{sum log p, count}`` on the device instead — the engine's
    fixed-order reduction writes ``sum log p`` straight into a slot (``log_prob_and_sum(x,
    sum_out=ring.slot(count))``), nothing else is launched per step — and reduces a whole bank of
    ``slots`` steps with ONE asynchronous ``all_reduce`` issued from a side stream when the bank
    is full (or on ``flush()``).  The compute stream never waits for a collective; ``means()``
    waits for the outstanding ones and returns the global mean NLL of every recorded step.
    """

    def __init__(self, device: torch.device | str, slots: int = 32, group: dist.ProcessGroup | None = None) -> None:
        self.device = torch.device(device)
        self.slots, self.group = int(slots), group
        self.sums = torch.zeros(2, self.slots, dtype=torch.float64, device=self.device)    # written by the engine
        self.counts = torch.zeros(2, self.slots, dtype=torch.float64, device=self.device)  # rewritten only when a count changes
        self._counts = [[None] * self.slots, [None] * self.slots]
        self.bank, self.i = 0, 0
        self._inflight: list = [None, None]  # per bank: (work | None, packet {sum, count} that is being reduced)
        self._done: list[Tensor] = []
        self.side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def _world(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def slot(self, count: int) -> Tensor:
        """The float64 element this step's ``sum log p`` goes into (pass it as ``sum_out``)."""
        if self.i == self.slots:
            self.flush()
        b, i = self.bank, self.i
        if self._inflight[b] is not None and i == 0:
            self._collect(b)  # the bank's previous reduction is retired before the bank is reused
        if self._counts[b][i] != count:  # constant across steps in practice: written once
            self.counts[b, i] = float(count)
            self._counts[b][i] = count
        self.i += 1
        return self.sums[b, i : i + 1]

    def flush(self) -> None:
        """Reduces the steps recorded since the last flush (asynchronously) and switches banks.  The
        collective works on a COPY {sum, count} of the bank (one small kernel per flush, not per step):
        the bank itself — in particular its count column — stays local and reusable."""
        n = self.i
        if n == 0:
            return
        b = self.bank
        work = None
        if self._world() > 1 and self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                packet = torch.stack((self.sums[b, :n], self.counts[b, :n]), dim=1)
                work = dist.all_reduce(packet, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            packet = torch.stack((self.sums[b, :n], self.counts[b, :n]), dim=1)
            if self._world() > 1:
                work = dist.all_reduce(packet, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight[b] = (work, packet)
        self.bank, self.i = 1 - b, 0

    def _collect(self, b: int) -> None:
        work, packet = self._inflight[b]
        if work is not None:
            work.wait()  # orders the current stream after the collective
        elif self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        if self.side is not None:
            packet.record_stream(torch.cuda.current_stream(self.device))
        self._done.append(-(packet[:, 0] / packet[:, 1]))
        self._inflight[b] = None

    def means(self) -> Tensor:
        """Global mean NLL of every step recorded so far (device float64 vector), in order."""
        self.flush()
        for b in (self.bank, 1 - self.bank):  # older bank first
            if self._inflight[b] is not None:
                self._collect(b)
        out = torch.cat(self._done) if self._done else torch.zeros(0, dtype=torch.float64, device=self.device)
        self._done = []
        return out


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
