"""zuko_b200 — B200-native (sm_100a) log-density and sampling engine for Zuko's flow hot path.

The public surface mirrors the reference (probabilists/zuko) for the hot path only:
``zuko_b200.flows.{MAF, NSF, NICE, ...}``, ``zuko_b200.lazy.Flow``,
``zuko_b200.distributions.NormalizingFlow`` — ``flow(c).log_prob(x)`` and
``flow(c).rsample(shape)`` run as hand-written CUDA kernels behind a C ABI
(``include/zuko_b200.h``).  There is no CPU fallback.
"""

__version__ = "0.1.0"

from . import distributions, flows, lazy, nn, transforms, utils  # noqa: F401
from .accel import AcceleratedFlow, accelerate  # noqa: F401
