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


class _FlowFunction(torch.autograd.Function):
    """Autograd seam of the engine: the forward is the usual engine call
    (``zk_flow_log_prob`` / ``zk_flow_forward``), the backward ONE ``zk_flow_backward`` call that
    recomputes the activations and returns d/d(x, c, every conditioner weight and bias, shared
    tables, rotation matrices) — what torch.autograd yields for the reference's eager graph."""

    @staticmethod
    def forward(ctx, call, mode, ldc, x2, c2, *params):  # noqa: ANN001
        # params = the call's conditioner / table parameters, then (log_prob mode) the base's loc, scale
        ctx.call, ctx.mode, ctx.ldc = call, mode, ldc
        ctx.has_c = c2 is not None
        ctx.generations = call._generations()
        ctx.save_for_backward(*([x2, c2] if c2 is not None else [x2]))
        lead = x2.shape[:-1]
        if mode == "log_prob":
            return call._run_log_prob(x2, c2, ldc, lead)
        return call._run_forward(x2, c2, ldc, lead)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):  # noqa: ANN001
        saved = ctx.saved_tensors
        x2 = saved[0]
        c2 = saved[1] if ctx.has_c else None
        need = ctx.needs_input_grad
        call = ctx.call
        call._check_generations(ctx.generations)
        n = len(call._params)
        gx, gc, pg = call._run_backward(ctx.mode, x2, c2, ctx.ldc, gouts, need[3], need[4], need[5 : 5 + n])
        base_g = []
        if len(need) > 5 + n:  # base loc / scale were passed: d log N(z; loc, scale) / d(loc, scale), z = f(x)
            base_g = [None, None]
            if need[5 + n] or need[6 + n]:
                z = call._run_forward(x2, c2, ctx.ldc, x2.shape[:-1])[0]
                base_g = list(call._base_param_grads(z, gouts[0], need[5 + n], need[6 + n]))
        return (None, None, None, gx, gc, *pg, *base_g)


class _FlowInverseFunction(torch.autograd.Function):
    """Autograd seam of the inverse direction (reparameterised sampling): forward =
    ``zk_flow_inverse`` (optionally with the log-density of the sample), backward = ONE
    ``zk_flow_inverse_backward`` call (implicit differentiation at the sample)."""

    @staticmethod
    def forward(ctx, call, with_lp, ldc, z2, c2, *params):  # noqa: ANN001
        ctx.call, ctx.with_lp, ctx.ldc = call, with_lp, ldc
        ctx.has_c = c2 is not None
        ctx.generations = call._generations()
        out = call._run_inverse(z2, c2, ldc, z2.shape[:-1], with_lp)
        x = out[0] if with_lp else out
        ctx.save_for_backward(*([x, z2, c2] if c2 is not None else [x, z2]))
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gouts):  # noqa: ANN001
        saved = ctx.saved_tensors
        x, z2 = saved[0], saved[1]
        c2 = saved[2] if ctx.has_c else None
        need = ctx.needs_input_grad
        call = ctx.call
        call._check_generations(ctx.generations)
        n = len(call._params)
        g_lp = gouts[1] if ctx.with_lp else None
        gz, gc, pg = call._run_inverse_backward(x, z2, c2, ctx.ldc, gouts[0], g_lp, need[3], need[4], need[5 : 5 + n])  # fmt: skip
        base_g = []
        if len(need) > 5 + n:  # explicit dependence of log p(x) = log N(z; loc, scale) + ... on loc, scale at fixed z
            base_g = [None, None]
            if g_lp is not None and (need[5 + n] or need[6 + n]):
                base_g = list(call._base_param_grads(z2, g_lp, need[5 + n], need[6 + n]))
        return (None, None, None, gz, gc, *pg, *base_g)


class FlowCall:
    """A ``zk_flow_desc`` over packed layers + DiagNormal / BoxUniform base, ready to be invoked."""

    def __init__(self, handles: Sequence, D: int, C: int, loc: Tensor | None, scale: Tensor | None,
                 sources: Sequence[dict] | None = None, keep: Sequence | None = None, base_kind: int = 0,
                 inverted: Sequence[bool] | None = None) -> None:  # fmt: skip
        self.D, self.C = D, C
        # members that are the INVERSE of their layer (LazyInverse, zuko/lazy.py:81-98); forward-only in the engine
        self._inverted = [bool(v) for v in inverted] if inverted is not None and any(inverted) else None
        self.base_kind = base_kind  # E.ZK_BASE_*: DiagNormal(loc, scale) or BoxUniform(lower=loc, upper=scale)
        self._handles = list(handles)
        # per layer: the tensors its parameter gradients belong to ({"weights", "biases"} of the
        # conditioner, {"phi"} shared table pieces, {"R"} rotation matrix) — see _FlowFunction
        self._sources = [dict(s) for s in sources] if sources is not None else [{} for _ in self._handles]
        self._keep = list(keep) if keep is not None else []  # owners of the zk_layer handles
        self._params: list[Tensor] = []
        self._slots: list[tuple[int, str, int]] = []  # (layer, kind, index) of every entry of _params
        for li, src in enumerate(self._sources):
            for kind in ("weights", "biases", "phi"):
                for k, t in enumerate(src.get(kind) or []):
                    if t is not None:
                        self._params.append(t)
                        self._slots.append((li, kind, k))
            if src.get("R") is not None:
                self._params.append(src["R"])
                self._slots.append((li, "R", 0))
        self._arr = (ctypes.c_void_p * max(1, len(handles)))(*[h.value if isinstance(h, ctypes.c_void_p) else h for h in handles])
        self._loc = None if loc is None else loc.detach().contiguous()
        self._scale = None if scale is None else scale.detach().contiguous()
        # a trainable DiagNormal base (UnconditionalDistribution(DiagNormal, loc, scale) with parameters,
        # zuko/lazy.py:242-287): loc / scale join the autograd seam as extra inputs
        self._base_src: list[Tensor] = []
        if base_kind == E.ZK_BASE_DIAG_NORMAL and loc is not None and (loc.requires_grad or scale.requires_grad):
            self._base_src = [loc, scale]
        if self._loc is not None:
            E.require_cuda(self._loc, "base loc")
            E.require_cuda(self._scale, "base scale")
        self._inv_arr = None if self._inverted is None else (ctypes.c_int * len(handles))(*[int(v) for v in self._inverted])
        self.desc = E.FlowDesc(len(handles), self._arr, D, C, _ptr(self._loc), _ptr(self._scale), base_kind, self._inv_arr)

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
        if self._wants_grad(x2, c2):
            z, ladj = _FlowFunction.apply(self, "forward", ldc, x2, c2, *self._params)
            return z.reshape(*lead, self.D), ladj.reshape(lead)
        return self._run_forward(x2, c2, ldc, lead)

    def _wants_grad(self, x2: Tensor, c2: Tensor | None) -> bool:
        if not torch.is_grad_enabled():
            return False
        if self._inverted is not None:  # forward-only members: callers route differentiable calls to the unfused chain
            return False
        return (x2.requires_grad or (c2 is not None and c2.requires_grad) or any(p.requires_grad for p in self._params)
                or bool(self._base_src))  # fmt: skip

    folded = None  # the same bijection with permutations folded into the autoregressive layers (forward-only), or None

    def _needs_grad(self, x: Tensor, c: Tensor | None) -> bool:
        if not torch.is_grad_enabled():
            return False
        return (x.requires_grad or (c is not None and c.requires_grad) or any(p.requires_grad for p in self._params)
                or bool(self._base_src))  # fmt: skip

    def best(self, x: Tensor, c: Tensor | None) -> "FlowCall":
        """The permutation-folded sibling when the call does not have to be differentiable, else this call."""
        if self.folded is not None and not self._needs_grad(x, c):
            return self.folded
        return self

    def usable(self, x: Tensor, c: Tensor | None, inverse_log_prob: bool = False) -> bool:
        """False when this call must go to the unfused chain instead: a flow with inverted members
        (LazyInverse) is forward-only in the engine — no autograd seam, no fused log-density of samples."""
        if self._inverted is None:
            return True
        if inverse_log_prob:
            return False
        if not torch.is_grad_enabled():
            return True
        return not (x.requires_grad or (c is not None and c.requires_grad) or any(p.requires_grad for p in self._params)
                    or bool(self._base_src))  # fmt: skip

    def _generations(self) -> tuple:
        return tuple(getattr(r, "generation", 0) for r in self._keep)

    def _check_generations(self, then: tuple) -> None:
        if then != self._generations():
            raise RuntimeError(
                "zuko_b200: the conditioner weights of this flow were refreshed in place (optimizer step) after the "
                "forward call of this graph; run backward() before the step, as torch requires for tensors modified in place"
            )

    def _base_param_grads(self, z: Tensor, g: Tensor, need_loc: bool, need_scale: bool):
        """d/d(loc, scale) of sum_b g_b log N(z_b; loc, scale) (torch/distributions/normal.py:87-102) at
        fixed z — reductions over the batch of element-wise terms (autograd plumbing of two D-vectors)."""
        loc, scale = self._base_src
        zf = z.detach().reshape(-1, self.D).to(torch.float32)
        gf = g.detach().reshape(-1, 1).to(torch.float32)
        u = (zf - loc.detach()) / scale.detach()
        g_loc = (gf * u / scale.detach()).sum(0).to(loc.dtype) if need_loc else None
        g_scale = (gf * (u * u - 1.0) / scale.detach()).sum(0).to(scale.dtype) if need_scale else None
        return g_loc, g_scale

    def _run_forward(self, x2: Tensor, c2: Tensor | None, ldc: int, lead) -> tuple[Tensor, Tensor]:
        x2, c2 = x2.detach(), (None if c2 is None else c2.detach())
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

    def log_prob(self, x: Tensor, c: Tensor | None, with_sum: bool = False, sum_out: Tensor | None = None):
        """``NormalizingFlow.log_prob(x)`` — zuko/distributions.py:115-119.  With
        ``with_sum`` also returns a device double holding ``sum(log_prob)`` (fixed-order
        reduction; the per-rank term of the mean NLL) — written into ``sum_out`` (one float64
        element on the device, e.g. a slot of ``dist.NllRing``) when given."""
        x2, c2, ldc, lead = _flatten(x, c if self.C else None, self.D)
        if self._wants_grad(x2, c2):
            lp = _FlowFunction.apply(self, "log_prob", ldc, x2, c2, *self._params, *self._base_src).reshape(lead)
            if not with_sum:
                return lp
            total = lp.detach().double().sum().reshape(1)
            if sum_out is not None:
                sum_out.copy_(total)
                total = sum_out
            return lp, total
        return self._run_log_prob(x2, c2, ldc, lead, with_sum, sum_out)

    def _run_log_prob(self, x2: Tensor, c2: Tensor | None, ldc: int, lead, with_sum: bool = False,
                      sum_out: Tensor | None = None):  # fmt: skip
        x2, c2 = x2.detach(), (None if c2 is None else c2.detach())
        B = x2.shape[0]
        lp = torch.empty(B, device=x2.device, dtype=torch.float32)
        total = None
        if with_sum:
            if sum_out is not None:
                if sum_out.dtype != torch.float64 or sum_out.numel() != 1 or sum_out.device != x2.device:
                    raise ValueError("sum_out must be one float64 element on the device of x")
                total = sum_out
            else:  # the reduction kernel overwrites its output: no fill needed
                total = torch.empty(1, device=x2.device, dtype=torch.float64)
        if B == 0:
            lp = lp.reshape(lead)
            if with_sum:
                total.zero_()
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
        zuko/distributions.py:129-138).  Differentiable w.r.t. z, c and the parameters."""
        z2, c2, ldc, lead = _flatten(z, c if self.C else None, self.D)
        if self._wants_grad(z2, c2):
            out = _FlowInverseFunction.apply(self, with_log_prob, ldc, z2, c2, *self._params,
                                             *(self._base_src if with_log_prob else []))
            if with_log_prob:
                return out[0].reshape(*lead, self.D), out[1].reshape(lead)
            return out.reshape(*lead, self.D)
        return self._run_inverse(z2, c2, ldc, lead, with_log_prob)

    def _run_inverse(self, z2: Tensor, c2: Tensor | None, ldc: int, lead, with_log_prob: bool = False):
        z2, c2 = z2.detach(), (None if c2 is None else c2.detach())
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

    def _run_backward(self, mode: str, x2: Tensor, c2: Tensor | None, ldc: int, gouts, need_x: bool, need_c: bool,
                      need_p: Sequence[bool]):  # fmt: skip
        """One ``zk_flow_backward`` call: returns (gx | None, gc | None, [param grads | None])."""
        L = E.lib()
        dev = x2.device
        B, D, C = x2.shape[0], self.D, self.C
        f32 = dict(device=dev, dtype=torch.float32)

        def dense(t):
            return None if t is None else t.detach().to(torch.float32).contiguous()

        g_lp = g_z = g_l = None
        if mode == "log_prob":
            g_lp = dense(gouts[0])
        else:
            g_z, g_l = dense(gouts[0]), dense(gouts[1])
        gx = torch.empty(B, D, **f32) if need_x else None
        gc = None
        if need_c and c2 is not None:
            gc = torch.zeros(C, **f32) if ldc == 0 else torch.empty(B, C, **f32)
        structs, pg, tables, keep = self._grad_structs(need_p, f32)
        if B:
            with torch.cuda.device(dev):
                want = L.zk_flow_backward_workspace_bytes(ctypes.byref(self.desc), B)
                minimum = L.zk_flow_backward_min_workspace_bytes(ctypes.byref(self.desc))
                floor = L.zk_flow_backward_workspace_bytes(ctypes.byref(self.desc), min(B, 1024))
                ws = E.Workspace.get(dev, want, max(minimum, floor))
                E.check(
                    L.zk_flow_backward(
                        ctypes.byref(self.desc), x2.data_ptr(), D, _ptr(c2), ldc, B, _ptr(g_z), D, _ptr(g_l), _ptr(g_lp),
                        _ptr(gx), D, _ptr(gc), C, structs, ws.data_ptr(), ws.numel(), E.stream_ptr(dev),
                    )
                )  # fmt: skip
        elif gx is not None:
            gx.zero_()
        self._split_tables(pg, tables, need_p)
        del keep
        return gx, gc, pg

    def _run_inverse_backward(self, x: Tensor, z2: Tensor, c2: Tensor | None, ldc: int, g_x, g_lp, need_z: bool,
                              need_c: bool, need_p: Sequence[bool]):  # fmt: skip
        """One ``zk_flow_inverse_backward`` call: returns (gz | None, gc | None, [param grads | None])."""
        L = E.lib()
        dev = x.device
        x = x.detach().reshape(-1, self.D).contiguous()
        B, D, C = x.shape[0], self.D, self.C
        f32 = dict(device=dev, dtype=torch.float32)
        g_x = None if g_x is None else g_x.detach().to(torch.float32).reshape(-1, D).contiguous()
        g_lp = None if g_lp is None else g_lp.detach().to(torch.float32).reshape(-1).contiguous()
        gz = torch.empty(B, D, **f32) if need_z else None
        gc = None
        if need_c and c2 is not None:
            gc = torch.zeros(C, **f32) if ldc == 0 else torch.empty(B, C, **f32)
        structs, pg, tables, keep = self._grad_structs(need_p, f32)
        if B:
            with torch.cuda.device(dev):
                want = L.zk_flow_inverse_backward_workspace_bytes(ctypes.byref(self.desc), B)
                minimum = L.zk_flow_inverse_backward_min_workspace_bytes(ctypes.byref(self.desc))
                floor = L.zk_flow_inverse_backward_workspace_bytes(ctypes.byref(self.desc), min(B, 1024))
                ws = E.Workspace.get(dev, want, max(minimum, floor))
                E.check(
                    L.zk_flow_inverse_backward(
                        ctypes.byref(self.desc), x.data_ptr(), D, _ptr(c2), ldc, B, _ptr(g_x), D, _ptr(g_lp),
                        z2.data_ptr(), D, _ptr(gz), D, _ptr(gc), C, structs, ws.data_ptr(), ws.numel(), E.stream_ptr(dev),
                    )
                )  # fmt: skip
        elif gz is not None:
            gz.zero_()
        self._split_tables(pg, tables, need_p)
        del keep
        return gz, gc, pg

    def _grad_structs(self, need_p: Sequence[bool], f32: dict):
        """Zero-initialised parameter-gradient buffers + the ``zk_layer_grads`` array pointing at them."""
        D = self.D
        # parameter gradient buffers (accumulated into by the engine => zero-initialised)
        nL = len(self._handles)
        pg: list[Tensor | None] = [None] * len(self._params)
        tables: dict[int, Tensor] = {}
        structs = (ctypes.POINTER(E.LayerGrads) * max(1, nL))()
        keep = []
        per_layer: dict[int, dict] = {}
        for i, (li, kind, k) in enumerate(self._slots):
            if not need_p[i]:
                continue
            d = per_layer.setdefault(li, {"weights": {}, "biases": {}, "phi": False, "R": None})
            p = self._params[i]
            if kind in ("weights", "biases"):
                pg[i] = torch.zeros(p.shape, **f32)
                d[kind][k] = pg[i]
            elif kind == "phi":
                d["phi"] = True
            else:
                pg[i] = torch.zeros(p.shape, **f32)
                d["R"] = pg[i]
        for li, d in per_layer.items():
            src = self._sources[li]
            lg = E.LayerGrads()
            n = len(src.get("weights") or [])
            if d["weights"] or d["biases"]:
                gw = (ctypes.c_void_p * n)(*[d["weights"][k].data_ptr() if k in d["weights"] else None for k in range(n)])
                gb = (ctypes.c_void_p * n)(*[d["biases"][k].data_ptr() if k in d["biases"] else None for k in range(n)])
                lg.grad_weight, lg.grad_bias = gw, gb
                keep += [gw, gb]
            if d["phi"]:
                P = sum(int(t[0].numel()) for t in src["phi"])
                tables[li] = torch.zeros(D, P, **f32)
                lg.grad_phi = tables[li].data_ptr()
            if d["R"] is not None:
                lg.grad_rotation = d["R"].data_ptr()
            keep.append(lg)
            structs[li] = ctypes.pointer(lg)
        return structs, pg, tables, keep

    def _split_tables(self, pg: list, tables: dict, need_p: Sequence[bool]) -> None:
        """Splits shared-table gradients back into the pieces of the ParameterList."""
        for i, (li, kind, k) in enumerate(self._slots):
            if kind == "phi" and need_p[i]:
                pieces = self._sources[li]["phi"]
                col = sum(int(t[0].numel()) for t in pieces[:k])
                w = int(pieces[k][0].numel())
                pg[i] = tables[li][:, col : col + w].reshape(pieces[k].shape).contiguous()

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
            want = L.zk_flow_host_workspace_bytes(ctypes.byref(self.desc), B) + (1 << 20)
            ws = E.Workspace.get(device, want, want)
            E.check(
                L.zk_flow_log_prob_host(
                    ctypes.byref(self.desc), x_host.data_ptr(), self.D, _ptr(c_host), ldc, B, out.data_ptr(),
                    ctypes.byref(total), ws.data_ptr(), ws.numel(), E.stream_ptr(device),
                )
            )  # fmt: skip
        return out, total.value
