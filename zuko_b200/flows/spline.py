"""Neural spline flows (zuko/flows/spline.py:21-62)."""

from __future__ import annotations

__all__ = ["NCSF", "NSF"]

from functools import partial

from math import pi

import torch

from ..distributions import BoxUniform
from ..lazy import UnconditionalDistribution
from ..transforms import CircularRQSTransform, MonotonicRQSTransform
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


class NCSF(MAF):
    """Neural circular spline flow (zuko/flows/spline.py:75-117): a :class:`MAF` whose univariate
    bijector is the circular RQS over ``[-pi, pi[`` and whose base is uniform over that box."""

    def __init__(self, features: int, context: int = 0, bins: int = 8, slope: float = 1e-3, **kwargs) -> None:
        super().__init__(
            features=features,
            context=context,
            univariate=partial(CircularRQSTransform, slope=slope),
            shapes=[(bins,), (bins,), (bins - 1,)],
            **kwargs,
        )
        self.base = UnconditionalDistribution(
            BoxUniform,
            lower=torch.full((features,), -pi - 1e-5),
            upper=torch.full((features,), pi + 1e-5),
            buffer=True,
        )
