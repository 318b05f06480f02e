"""Flow zoo of the hot path (zuko/flows/__init__.py, restricted to SURVEY §8)."""

from .autoregressive import MAF, MaskedAutoregressiveTransform
from .coupling import NICE, GeneralCouplingTransform, RealNVP
from .elementwise import ElementWiseTransform
from .spline import NCSF, NSF
from ..lazy import Flow, UnconditionalDistribution, UnconditionalTransform

__all__ = [
    "MAF",
    "NCSF",
    "NICE",
    "NSF",
    "ElementWiseTransform",
    "Flow",
    "GeneralCouplingTransform",
    "MaskedAutoregressiveTransform",
    "RealNVP",
    "UnconditionalDistribution",
    "UnconditionalTransform",
]
