"""Shape glue of the hot path: ``Partial``, ``broadcast`` and ``unpack``.

Host-side mirror of zuko/utils.py:26-115 (Partial), :212-244 (broadcast) and :596-622
(unpack).  Nothing here touches the device.
"""

from __future__ import annotations

__all__ = ["Partial", "broadcast", "unpack"]

import math
from collections.abc import Callable, Sequence
from typing import Any

import torch
import torch.nn as nn
from torch import Size, Tensor


class Partial(nn.Module):
    """``functools.partial`` as an ``nn.Module``: tensor arguments of ``f`` become buffers
    (``buffer=True``) or parameters, so ``state_dict`` / ``.to()`` see them.

    Positional tensors are registered as ``_0, _1, ...`` and keyword tensors under their
    own name — the state-dict layout of zuko/utils.py:60-86 (e.g. ``base.loc``,
    ``base.scale`` for the DiagNormal base of MAF/NSF).
    """

    def __init__(self, f: Callable, /, *args, buffer: bool = False, **kwargs) -> None:
        super().__init__()
        self.f = f
        self._nargs = len(args)
        self._keys = list(kwargs)
        named = [(f"_{i}", a) for i, a in enumerate(args)] + list(kwargs.items())
        for name, value in named:
            if not torch.is_tensor(value):
                setattr(self, name, value)
            elif buffer:
                self.register_buffer(name, value)
            else:
                self.register_parameter(name, nn.Parameter(value))

    @property
    def args(self) -> Sequence[Any]:
        return [getattr(self, f"_{i}") for i in range(self._nargs)]

    @property
    def kwargs(self) -> dict[str, Any]:
        return {k: getattr(self, k) for k in self._keys}

    def extra_repr(self) -> str:
        return "" if isinstance(self.f, nn.Module) else f"(f): {self.f}"

    def forward(self, *args, **kwargs) -> Any:  # noqa: ANN401
        return self.f(*self.args, *args, **self.kwargs, **kwargs)


def broadcast(*tensors: Tensor, ignore: int | Sequence[int] = 0) -> list[Tensor]:
    """Broadcasts the leading dimensions of ``tensors``, leaving the last ``ignore``
    dimensions of each untouched (zuko/utils.py:212-244)."""
    if isinstance(ignore, int):
        ignore = [ignore] * len(tensors)
    heads = [t.shape[: t.dim() - k] for t, k in zip(tensors, ignore, strict=True)]
    common = torch.broadcast_shapes(*heads)
    return [
        torch.broadcast_to(t, common + t.shape[t.dim() - k :])
        for t, k in zip(tensors, ignore, strict=True)
    ]


def unpack(x: Tensor, shapes: Sequence[Size]) -> Sequence[Tensor]:
    """Splits the last dimension of ``x`` into tensors of trailing shapes ``shapes``
    (zuko/utils.py:596-622)."""
    sizes = [math.prod(s) for s in shapes]
    parts = x.split(sizes, -1)
    return tuple(p.reshape(*p.shape[:-1], *s) for p, s in zip(parts, shapes, strict=True))
