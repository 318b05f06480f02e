"""Masked autoregressive transformations and flows.

Host-side mirror of zuko/flows/autoregressive.py: ``MaskedAutoregressiveTransform``
(:24-218) and ``MAF`` (:221-316).  Construction (orders, adjacency, masks, parameter
initialisation) reproduces the reference bit for bit; evaluation is handed to the B200
engine as one packed ``zk_layer`` (conditioner GEMMs + fused bijector/ladj kernel).
"""

from __future__ import annotations

__all__ = ["MAF", "MaskedAutoregressiveTransform"]

from collections.abc import Callable, Sequence
from math import ceil

import torch
from torch import BoolTensor, LongTensor, Size, Tensor
from torch.distributions import Transform

from .. import _engine as E
from ..distributions import DiagNormal
from ..lazy import Flow, LazyTransform, UnconditionalDistribution
from ..nn import MaskedMLP
from ..transforms import AutoregressiveTransform, MonotonicAffineTransform
from ._packed import PackedLayerMixin, resolve_univariate, total_of
from .elementwise import ElementWiseTransform


def dag_diameter(adjacency: BoolTensor) -> int:
    """Number of topological generations of the DAG ``adjacency`` (``[i, j]``: i depends
    on j) = sequential passes the inverse needs (zuko/flows/autoregressive.py:154-185).
    Raises if the graph has a cycle."""
    n = adjacency.shape[0]
    pending = adjacency.sum(dim=1).tolist()  # unresolved dependencies per node
    frontier = [i for i in range(n) if pending[i] == 0]
    generations = 0
    resolved = 0
    while frontier:
        generations += 1
        resolved += len(frontier)
        nxt = []
        for node in frontier:
            for child in adjacency[:, node].nonzero().flatten().tolist():
                pending[child] -= 1
                if pending[child] == 0:
                    nxt.append(child)
        frontier = nxt
    assert resolved == n, "The graph contains cycles."
    return generations


class MaskedAutoregressiveTransform(PackedLayerMixin, LazyTransform):
    """Lazy masked autoregressive transformation ``y_i = f(x_i | x_<i, c)``.

    Arguments follow zuko/flows/autoregressive.py:88-99: ``features``, ``context``,
    ``passes`` (None = fully autoregressive, 2 = coupling), ``order``, ``adjacency``
    (custom dependency DAG; overrides ``order`` / ``passes``), ``univariate`` / ``shapes``
    (engine bijectors: ``MonotonicAffineTransform``, ``MonotonicRQSTransform``) and
    ``**kwargs`` for :class:`zuko_b200.nn.MaskedMLP`.
    """

    def __new__(cls, features: int | None = None, context: int = 0, passes: int | None = None,
                order: LongTensor | None = None, adjacency: BoolTensor | None = None, *args, **kwargs) -> LazyTransform:  # fmt: skip
        if features is None or features > 1:
            return super().__new__(cls)
        return ElementWiseTransform(features, context, *args, **kwargs)  # single feature: nothing to mask

    def __init__(
        self,
        features: int,
        context: int = 0,
        passes: int | None = None,
        order: LongTensor | None = None,
        adjacency: BoolTensor | None = None,
        univariate: Callable[..., Transform] = MonotonicAffineTransform,
        shapes: Sequence[Size] = ((), ()),
        **kwargs,
    ) -> None:
        super().__init__()
        self.univariate, self.shapes = univariate, shapes
        self.total = total_of(shapes)
        self._uni = resolve_univariate(univariate, shapes)
        self.features, self.context = features, context
        self.register_buffer("order", None)

        ctx_cols = None
        if adjacency is None:
            passes = features if passes is None else passes
            order = torch.arange(features) if order is None else torch.as_tensor(order, dtype=int)
            assert order.ndim == 1, "'order' should be a vector."
            assert order.shape[0] == features, f"'order' should have {features} elements."
            self.passes = min(max(passes, 1), features)
            # features sharing an order class are transformed in the same pass
            self.order = torch.div(order, ceil(features / self.passes), rounding_mode="floor")
            deps = self.order[:, None] > self.order
        else:
            adjacency = torch.as_tensor(adjacency, dtype=bool)
            assert adjacency.ndim == 2, "'adjacency' should be a matrix."
            assert adjacency.shape[0] == features, f"'adjacency' should have {features} rows."
            assert adjacency.shape[1] in (features, features + context), (
                f"'adjacency' should have {features} or {features + context} columns."
            )
            if adjacency.shape[1] > features:
                ctx_cols = adjacency[:, features:]
            deps = adjacency[:, :features]
            assert deps.diag().all(), "'adjacency' should have ones on the diagonal."
            deps = deps * ~torch.eye(features, dtype=bool)
            self.passes = dag_diameter(deps)

        if context > 0:
            if ctx_cols is None:
                ctx_cols = torch.ones((features, context), dtype=bool)
            deps = torch.cat((deps, ctx_cols), dim=1)

        # one conditioner output row per (feature, parameter): row index d * total + p
        self.hyper = MaskedMLP(torch.repeat_interleave(deps, repeats=self.total, dim=0), **kwargs)

    def extra_repr(self) -> str:
        if self.order is None:
            return f"(base): {self._describe_base()}\n(passes): {self.passes}"
        order = self.order.tolist()
        if len(order) > 10:
            order = str(order[:5] + [...] + order[-5:]).replace("Ellipsis", "...")
        return f"(base): {self._describe_base()}\n(order): {order}"

    def _layer_desc(self):
        hyper, keep = self.hyper.mlp_desc()
        desc = self._base_desc(E.ZK_LAYER_AUTOREGRESSIVE)
        import ctypes

        desc.hyper = ctypes.pointer(hyper)
        keep = [hyper, keep]
        if self.order is not None:  # order classes let the engine run the dimension-sequential inverse
            host = self.order.detach().to("cpu", torch.int64).tolist()
            arr = (ctypes.c_int64 * self.features)(*host)
            desc.order = arr
            keep.append(arr)
        return desc, keep

    def forward(self, c: Tensor | None = None) -> Transform:
        return AutoregressiveTransform(self, c)


class MAF(Flow):
    """Masked autoregressive flow (zuko/flows/autoregressive.py:221-316): ``transforms``
    autoregressive layers whose feature order alternates ascending / descending (or is
    random with ``randperm=True``), over a standard ``DiagNormal`` base."""

    def __init__(self, features: int, context: int = 0, transforms: int = 3, randperm: bool = False, **kwargs) -> None:
        ascending = torch.arange(features)
        layers = []
        for i in range(transforms):
            if randperm:
                order = torch.randperm(features)
            else:
                order = ascending if i % 2 == 0 else ascending.flip(0)
            layers.append(MaskedAutoregressiveTransform(features=features, context=context, order=order, **kwargs))
        base = UnconditionalDistribution(DiagNormal, loc=torch.zeros(features), scale=torch.ones(features), buffer=True)
        super().__init__(layers, base)
