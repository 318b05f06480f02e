"""Tensor-level wrappers of the C ABI: shape normalisation, workspace, stream.

Calling conventions honoured: ``x`` is ``(*, D)``; ``c`` is ``None``, ``(C,)``
or ``(*, C)``; leading dimensions broadcast (zuko/flows/autoregressive.py:209,
zuko/utils.py:236-244) and the result carries the broadcast batch shape.  Inputs are
taken as contiguous fp32 copies when they are not already.
"""

from __future__ import annotations

import ctypes
from collections.abc import Sequence

import torch
from torch import Tensor

from . import _engine as E


def _flatten(x: Tensor, c: Tensor | None, D: int):
    """Returns (x2 (B, D), c2 or None, ldc, lead_shape)."""
    E.require_cuda(x, "input")
    if x.requires_grad or (c is not None and c.requires_grad):
        raise NotImplementedError(
            "zuko_b200: inputs that require grad are not supported (the engine is forward-only; "
            "the backward pass is listed as next in SURVEY §8f)"
        )
    if x.shape[-1] != D:
        raise ValueError(f"zuko_b200: expected {D} features in the last dimension, got {tuple(x.shape)}")
    lead = x.shape[:-1]
    ldc = 0
    c2 = None
    if c is not None:
        E.require_cuda(c, "context")
        if c.device != x.device:
            raise E.EngineError("zuko_b200: input and context live on different devices")
        if c.dim() == 1:
            c2 = c.contiguous()
        else:
            lead = torch.broadcast_shapes(lead, c.shape[:-1])
            c2 = c.expand(*lead, c.shape[-1]).reshape(-1, c.shape[-1]).contiguous()
            ldc = c2.shape[-1]
    x2 = x.expand(*lead, D).reshape(-1, D).contiguous()
    return x2, c2, ldc, lead


def _ptr(t: Tensor | None):
    return None if t is None else t.data_ptr()


def layer_forward(handle, D: int, x: Tensor, c: Tensor | None) -> tuple[Tensor, Tensor]:
    """``t(c).call_and_ladj(x)`` of one packed layer: returns ``y (*, D)``, ``ladj (*)``."""
    x2, c2, ldc, lead = _flatten(x, c, D)
    B = x2.shape[0]
    y = torch.empty_like(x2)
    ladj = torch.empty(B, device=x2.device, dtype=torch.float32)
    if B == 0:
        return y.reshape(*lead, D), ladj.reshape(lead)
    L = E.lib()
    with torch.cuda.device(x2.device):
        need = L.zk_layer_workspace_bytes(handle, B)
        ws = E.Workspace.get(x2.device, need, need) if need else None
        E.check(
            L.zk_layer_forward(
                handle, x2.data_ptr(), D, _ptr(c2), ldc, B, y.data_ptr(), D, ladj.data_ptr(), 0,
                _ptr(ws), ws.numel() if ws is not None else 0, E.stream_ptr(x2.device),
            )
        )  # fmt: skip
    return y.reshape(*lead, D), ladj.reshape(lead)


def layer_inverse(handle, D: int, y: Tensor, c: Tensor | None) -> Tensor:
    """``t(c).inv(y)`` of one packed layer."""
    y2, c2, ldc, lead = _flatten(y, c, D)
    B = y2.shape[0]
    x = torch.empty_like(y2)
    if B == 0:
        return x.reshape(*lead, D)
    L = E.lib()
    with torch.cuda.device(y2.device):
        need = L.zk_layer_workspace_bytes(handle, B)
        ws = E.Workspace.get(y2.device, need, need) if need else None
        E.check(
            L.zk_layer_inverse(
                handle, y2.data_ptr(), D, _ptr(c2), ldc, B, x.data_ptr(), D, _ptr(ws),
                ws.numel() if ws is not None else 0, E.stream_ptr(y2.device),
            )
        )  # fmt: skip
    return x.reshape(*lead, D)


class FlowCall:
    """A ``zk_flow_desc`` over packed layers + DiagNormal base, ready to be invoked."""

    def __init__(self, handles: Sequence, D: int, C: int, loc: Tensor | None, scale: Tensor | None) -> None:
        self.D, self.C = D, C
        self._handles = list(handles)
        self._arr = (ctypes.c_void_p * max(1, len(handles)))(*[h.value if isinstance(h, ctypes.c_void_p) else h for h in handles])
        self._loc = None if loc is None else loc.detach().contiguous()
        self._scale = None if scale is None else scale.detach().contiguous()
        if self._loc is not None:
            E.require_cuda(self._loc, "base loc")
            E.require_cuda(self._scale, "base scale")
        self.desc = E.FlowDesc(len(handles), self._arr, D, C, _ptr(self._loc), _ptr(self._scale))

    def _ws(self, device, B: int):
        L = E.lib()
        want = L.zk_flow_workspace_bytes(ctypes.byref(self.desc), B)
        minimum = L.zk_flow_min_workspace_bytes(ctypes.byref(self.desc))
        # never below a 4096-row chunk when the batch is that large
        floor = L.zk_flow_workspace_bytes(ctypes.byref(self.desc), min(B, 4096))
        return E.Workspace.get(device, want, max(minimum, floor))

    def forward(self, x: Tensor, c: Tensor | None) -> tuple[Tensor, Tensor]:
        """``transform.call_and_ladj(x)`` — zuko/transforms.py:141-150."""
        x2, c2, ldc, lead = _flatten(x, c if self.C else None, self.D)
        B = x2.shape[0]
        z = torch.empty_like(x2)
        ladj = torch.empty(B, device=x2.device, dtype=torch.float32)
        if B:
            with torch.cuda.device(x2.device):
                ws = self._ws(x2.device, B)
                E.check(
                    E.lib().zk_flow_forward(
                        ctypes.byref(self.desc), x2.data_ptr(), self.D, _ptr(c2), ldc, B, z.data_ptr(),
                        self.D, ladj.data_ptr(), ws.data_ptr(), ws.numel(), E.stream_ptr(x2.device),
                    )
                )  # fmt: skip
        return z.reshape(*lead, self.D), ladj.reshape(lead)

    def log_prob(self, x: Tensor, c: Tensor | None, with_sum: bool = False):
        """``NormalizingFlow.log_prob(x)`` — zuko/distributions.py:115-119.  With
        ``with_sum`` also returns a device double holding ``sum(log_prob)`` (fixed-order
        reduction; the per-rank term of the mean NLL)."""
        x2, c2, ldc, lead = _flatten(x, c if self.C else None, self.D)
        B = x2.shape[0]
        lp = torch.empty(B, device=x2.device, dtype=torch.float32)
        total = torch.zeros(1, device=x2.device, dtype=torch.float64) if with_sum else None
        if B == 0:
            lp = lp.reshape(lead)
            return (lp, total) if with_sum else lp
        with torch.cuda.device(x2.device):
            ws = self._ws(x2.device, max(B, 1))
            E.check(
                E.lib().zk_flow_log_prob(
                    ctypes.byref(self.desc), x2.data_ptr(), self.D, _ptr(c2), ldc, B, lp.data_ptr(),
                    _ptr(total), ws.data_ptr(), ws.numel(), E.stream_ptr(x2.device),
                )
            )  # fmt: skip
        lp = lp.reshape(lead)
        return (lp, total) if with_sum else lp

    def inverse(self, z: Tensor, c: Tensor | None, with_log_prob: bool = False):
        """``transform.inv(z)`` (and the log-density of the result when ``with_log_prob``,
        zuko/distributions.py:129-138)."""
        z2, c2, ldc, lead = _flatten(z, c if self.C else None, self.D)
        B = z2.shape[0]
        x = torch.empty_like(z2)
        lp = torch.empty(B, device=z2.device, dtype=torch.float32) if with_log_prob else None
        if B:
            with torch.cuda.device(z2.device):
                ws = self._ws(z2.device, B)
                E.check(
                    E.lib().zk_flow_inverse(
                        ctypes.byref(self.desc), z2.data_ptr(), self.D, _ptr(c2), ldc, B, x.data_ptr(),
                        self.D, _ptr(lp), ws.data_ptr(), ws.numel(), E.stream_ptr(z2.device),
                    )
                )  # fmt: skip
        x = x.reshape(*lead, self.D)
        return (x, lp.reshape(lead)) if with_log_prob else x

    def log_prob_host(self, x_host: Tensor, c_host: Tensor | None, device: torch.device, out: Tensor | None = None):
        """End-to-end entry: HOST (pinned) inputs, HOST output; H2D / compute / D2H are
        pipelined over row chunks inside ``zk_flow_log_prob_host``.  Returns
        ``(log_prob_host, sum_log_prob)``; synchronous."""
        assert not x_host.is_cuda and x_host.dtype == torch.float32 and x_host.dim() == 2
        x_host = x_host.contiguous()
        B = x_host.shape[0]
        ldc = 0
        if c_host is not None and self.C:
            assert not c_host.is_cuda and c_host.dtype == torch.float32
            c_host = c_host.contiguous()
            ldc = 0 if c_host.dim() == 1 else self.C
        else:
            c_host = None
        if out is None:
            out = torch.empty(B, dtype=torch.float32, pin_memory=True)
        total = ctypes.c_double(0.0)
        with torch.cuda.device(device):
            L = E.lib()
            chunk = max(4096, -(-B // 8))
            want = L.zk_flow_workspace_bytes(ctypes.byref(self.desc), chunk) + 2 * chunk * (self.D + self.C + 1) * 4 + (1 << 20)
            ws = E.Workspace.get(device, want, want)
            E.check(
                L.zk_flow_log_prob_host(
                    ctypes.byref(self.desc), x_host.data_ptr(), self.D, _ptr(c_host), ldc, B, out.data_ptr(),
                    ctypes.byref(total), ws.data_ptr(), ws.numel(), E.stream_ptr(device),
                )
            )  # fmt: skip
        return out, total.value
