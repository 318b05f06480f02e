"""Neural spline flows (zuko/flows/spline.py:21-62)."""

from __future__ import annotations

__all__ = ["NSF"]

from functools import partial

from ..transforms import MonotonicRQSTransform
from .autoregressive import MAF


class NSF(MAF):
    """Neural spline flow: a :class:`MAF` whose univariate bijector is the monotonic
    rational-quadratic spline with ``bins`` bins over [-5, 5] (features outside pass
    through unchanged).  ``passes=2`` gives coupling layers."""

    def __init__(self, features: int, context: int = 0, bins: int = 8, slope: float = 1e-3, **kwargs) -> None:
        super().__init__(
            features=features,
            context=context,
            univariate=partial(MonotonicRQSTransform, slope=slope),
            shapes=[(bins,), (bins,), (bins - 1,)],
            **kwargs,
        )
