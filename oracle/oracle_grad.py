"""ORACLE (gradients) — TEST INFRASTRUCTURE ONLY.

CPU restatement, in fp64 numpy, of what ``torch.autograd`` computes when the reference's
training loop calls ``loss = -flow(c).log_prob(x).mean(); loss.backward()``
(README.md:43-49, tests/test_flows.py:22-29): a hand-written reverse-mode walk over the SAME
op sequence the reference executes in its forward pass

* MonotonicRQSTransform.__init__ / bin / call_and_ladj   zuko/transforms.py:469-567
* MonotonicAffineTransform                               zuko/transforms.py:426-446
* SoftclipTransform                                      zuko/transforms.py:299-316
* MaskedLinear / MLP with ReLU                           zuko/nn.py:217-218, 122-192
* Autoregressive / Coupling / DependentTransform         zuko/transforms.py:966-1073, 210-214
* ComposedTransform + NormalizingFlow.log_prob           zuko/transforms.py:141-150,
                                                         zuko/distributions.py:115-119

Parity status: PINNED — ``tests/test_oracle_grad.py`` checks every function here against
``tests/golden/grad_*.npz`` (gradients produced by torch.autograd on the unmodified reference
in fp64, ``tests/golden/make_golden_grad.py``).

The formulation is deliberately the reference's (knots by cumsum, bin limits gathered as
x0/x1/y0/y1, ``where(mask, ...)``), not the engine's (bin widths straight from the softmax
numerators), so that the two derivations check each other.

Only ``tests/`` and the CPU legs of ``bench.py`` may import this module.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import oracle as O

F = np.float64


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #


def _softclip(v, a):
    """v / (1 + |v| a) and its derivative 1 / (1 + |v| a)^2 (transforms.py:480-482, 436)."""
    den = 1.0 + np.abs(v) * a
    return v / den, 1.0 / (den * den)


def _softmax(v):
    e = np.exp(v - v.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


# --------------------------------------------------------------------------- #
# univariate bijectors: (gy, gl) -> (gx, gphi); all arrays (B, D[, P])
# --------------------------------------------------------------------------- #


def rqs_backward(x, phi, gy, gl, bins: int, bound=5.0, slope=1e-3):
    """Reverse mode of MonotonicRQSTransform(*unpack(phi)).call_and_ladj(x).

    x (B, D); phi (B, D, 3K-1) or shared (D, 3K-1); gy = dL/dy (B, D); gl = dL/dladj
    PER ELEMENT (B, D).  Returns (gx (B, D), gphi with the shape of phi — summed over the
    batch when phi is shared)."""
    K = bins
    x = np.asarray(x, F)
    B, D = x.shape
    phi_in = np.asarray(phi, F)
    shared = phi_in.ndim == 2
    phi = np.broadcast_to(phi_in, (B, D, 3 * K - 1))
    gy = np.asarray(gy, F)
    gl = np.asarray(gl, F)
    absL = abs(np.log(slope))
    w, h, d = phi[..., :K], phi[..., K : 2 * K], phi[..., 2 * K :]
    # ---- forward of __init__ (transforms.py:480-490)
    cw, dcw = _softclip(w, 2.0 / absL)
    ch, dch = _softclip(h, 2.0 / absL)
    cd, dcd = _softclip(d, 1.0 / absL)
    sw, sh = _softmax(cw), _softmax(ch)
    zeros = np.zeros((B, D, 1), F)
    X = bound * (2.0 * np.cumsum(np.concatenate([zeros, sw], -1), -1) - 1.0)
    Y = bound * (2.0 * np.cumsum(np.concatenate([zeros, sh], -1), -1) - 1.0)
    Dv = np.exp(np.concatenate([zeros, cd, zeros], -1))
    # ---- bin (transforms.py:499-523, 555)
    k = (X < x[..., None]).sum(-1) - 1
    mask = (k >= 0) & (k < K)
    km = k % K
    take = lambda a, i: np.take_along_axis(a, i[..., None], -1)[..., 0]  # noqa: E731
    x0, x1 = take(X, km), take(X, km + 1)
    y0, y1 = take(Y, km), take(Y, km + 1)
    d0, d1 = take(Dv, km), take(Dv, km + 1)
    dx, dy = x1 - x0, y1 - y0
    s = dy / dx
    z = mask * (x - x0) / dx
    # ---- call_and_ladj (transforms.py:556-567)
    q = z * (1.0 - z)
    t = d0 + d1 - 2.0 * s
    num = s * z * z + d0 * q
    den = s + t * q
    m = 2.0 * s * q + d0 * (1.0 - z) ** 2 + d1 * z * z
    # y = y0 + dy * num / den ; ladj = log(s^2 m / den^2)
    # ---- reverse: partials at fixed (y0, dy, s, z, d0, d1)
    gyv = np.where(mask, gy, 0.0)
    glv = np.where(mask, gl, 0.0)
    r = num / den
    dy_dz = dy * ((2.0 * s * z + d0 * (1.0 - 2.0 * z)) * den - num * t * (1.0 - 2.0 * z)) / den**2
    dy_ds = dy * (z * z * den - num * (1.0 - 2.0 * q)) / den**2
    dy_dd0 = dy * (q * den - num * q) / den**2
    dy_dd1 = dy * (-num * q) / den**2
    dl_ds = 2.0 / s + 2.0 * q / m - 2.0 * (1.0 - 2.0 * q) / den
    dl_dz = (2.0 * s * (1.0 - 2.0 * z) - 2.0 * d0 * (1.0 - z) + 2.0 * d1 * z) / m - 2.0 * t * (1.0 - 2.0 * z) / den
    dl_dd0 = (1.0 - z) ** 2 / m - 2.0 * q / den
    dl_dd1 = z * z / m - 2.0 * q / den
    Gs = gyv * dy_ds + glv * dl_ds
    Gz = gyv * dy_dz + glv * dl_dz
    Gd0 = gyv * dy_dd0 + glv * dl_dd0
    Gd1 = gyv * dy_dd1 + glv * dl_dd1
    Gdy = gyv * r + Gs / dx  # s = dy / dx
    Gdx = -Gs * s / dx - Gz * z / dx  # z = (x - x0) / dx
    gx = np.where(mask, Gz / dx, gy)  # outside the domain: identity (transforms.py:567)
    Gx0 = -Gz / dx - Gdx  # dx = x1 - x0
    Gx1 = Gdx
    Gy0 = gyv - Gdy
    Gy1 = Gdy
    # ---- scatter into the knot tables
    gX = np.zeros_like(X)
    gY = np.zeros_like(Y)
    gDv = np.zeros_like(Dv)

    def put(tbl, i, v):
        np.put_along_axis(tbl, i[..., None], take(tbl, i)[..., None] + v[..., None], -1)

    put(gX, km, Gx0)
    put(gX, km + 1, Gx1)
    put(gY, km, Gy0)
    put(gY, km + 1, Gy1)
    put(gDv, km, Gd0)
    put(gDv, km + 1, Gd1)

    # ---- reverse of __init__
    def knots_back(gK, sm, dclip):
        gcum = 2.0 * bound * gK  # knots = bound * (2 cumsum - 1)
        gpad = np.cumsum(gcum[..., ::-1], -1)[..., ::-1]  # reverse of cumsum
        gsm = gpad[..., 1:]  # drop the left pad
        gc = sm * (gsm - (sm * gsm).sum(-1, keepdims=True))  # softmax
        return gc * dclip

    g_w = knots_back(gX, sw, dcw)
    g_h = knots_back(gY, sh, dch)
    g_d = (gDv * Dv)[..., 1:-1] * dcd
    gphi = np.concatenate([g_w, g_h, g_d], -1)
    if shared:
        gphi = gphi.sum(0)
    return gx, gphi


def affine_backward(x, phi, gy, gl, slope=1e-3):
    """Reverse mode of MonotonicAffineTransform(shift, scale).call_and_ladj(x)
    (transforms.py:435-446); phi[..., 0] = shift, phi[..., 1] = unconstrained log-scale."""
    x = np.asarray(x, F)
    B, D = x.shape
    phi_in = np.asarray(phi, F)
    shared = phi_in.ndim == 2
    phi = np.broadcast_to(phi_in, (B, D, 2))
    ls, dls = _softclip(phi[..., 1], 1.0 / abs(np.log(slope)))
    e = np.exp(ls)
    gx = gy * e
    g_shift = gy
    g_a = (gy * x * e + gl) * dls
    gphi = np.stack([g_shift, g_a], -1)
    if shared:
        gphi = gphi.sum(0)
    return gx, gphi


def softclip_backward(x, gy, gl, bound):
    """SoftclipTransform: y = x / (1 + |x/B|), ladj = -2 log1p(|x/B|) (transforms.py:309-316)."""
    x = np.asarray(x, F)
    t = np.abs(x) / bound
    return gy / (1.0 + t) ** 2 - 2.0 * gl * np.sign(x) / (bound * (1.0 + t))


# --------------------------------------------------------------------------- #
# conditioner
# --------------------------------------------------------------------------- #


@dataclass
class CondGrads:
    weights: list = field(default_factory=list)
    biases: list = field(default_factory=list)


def conditioner_backward(cond: O.Conditioner, inp, gout):
    """Reverse mode of the (masked) MLP (nn.py:217-218, 311-313), residual blocks included
    (nn.py:195-199).  inp (B, in), gout (B, out).  Returns (ginp, CondGrads) — weight gradients are
    w.r.t. the RAW weights, i.e. already multiplied by the mask."""
    n = len(cond.weights)
    acts, res = cond.flags()
    a, outs = cond.trace(inp)
    g = np.asarray(gout, F)  # dL/d(output of linear layer i, before its activation)
    out = CondGrads([None] * n, [None] * n)
    pending = {}  # index of a layer input -> gradient arriving through a residual connection
    for i in reversed(range(n)):
        mk = 1.0 if cond.masks[i] is None else cond.masks[i]
        out.weights[i] = (g.T @ a[i]) * mk
        out.biases[i] = None if cond.biases[i] is None else g.sum(0)
        ga = g @ (cond.weights[i] * mk) + pending.pop(i, 0.0)
        if res[i]:
            pending[i - 1] = g
        if i > 0:
            g = ga * O.ACTIVATIONS[acts[i - 1]][1](outs[i - 1]) if acts[i - 1] else ga
        else:
            g = ga
    return g, out


def _cond_forward(cond: O.Conditioner, inp):
    return cond.trace(inp)[0][-1]


# --------------------------------------------------------------------------- #
# layers and flows
# --------------------------------------------------------------------------- #


@dataclass
class LayerGrads:
    hyper: CondGrads | None = None
    phi: np.ndarray | None = None  # shared (D, P) table
    R: np.ndarray | None = None  # rotation matrix


def _ctx_rows(c, B):
    if c is None:
        return None
    c = np.asarray(c, F)
    return np.broadcast_to(c, (B, c.shape[-1])) if c.ndim == 1 else c


def _uni_backward(layer: O.Layer, x, phi, gy, gl_row):
    gl = np.broadcast_to(np.asarray(gl_row, F)[:, None], x.shape)  # ladj.sum(-1): transforms.py:210-214
    if layer.univariate == "rqs":
        return rqs_backward(x, phi, gy, gl, layer.bins, layer.bound, layer.slope)
    if layer.univariate == "crqs":  # the circular shift is a translation almost everywhere: d/dx = 1
        return rqs_backward(O.circular_shift(x, layer.bound), phi, gy, gl, layer.bins, layer.bound, layer.slope)
    return affine_backward(x, phi, gy, gl, layer.slope)


def layer_backward(layer: O.Layer, x, c, gy, gl):
    """Reverse mode of ``layer.forward`` (oracle.Layer.forward): given gy (B, D) and gl (B)
    returns (gx (B, D), gc (B, C) or None, LayerGrads)."""
    x = np.asarray(x, F)
    gy = np.asarray(gy, F)
    gl = np.asarray(gl, F)
    B, D = x.shape
    cc = _ctx_rows(c, B)
    P = 3 * layer.bins - 1 if layer.univariate in ("rqs", "crqs") else 2
    if layer.kind == "autoregressive":
        inp = x if cc is None else np.concatenate([x, cc], -1)  # flows/autoregressive.py:209
        phi = _cond_forward(layer.hyper, inp).reshape(B, D, P)
        gx, gphi = _uni_backward(layer, x, phi, gy, gl)
        ginp, cg = conditioner_backward(layer.hyper, inp, gphi.reshape(B, D * P))
        gx = gx + ginp[:, :D]
        return gx, (None if cc is None else ginp[:, D:]), LayerGrads(hyper=cg)
    if layer.kind == "coupling":
        ia, ib = np.nonzero(layer.mask)[0], np.nonzero(~layer.mask)[0]
        xa, xb = x[:, ia], x[:, ib]
        inp = xa if cc is None else np.concatenate([xa, cc], -1)
        phi = _cond_forward(layer.hyper, inp).reshape(B, len(ib), P)
        gxb, gphi = _uni_backward(layer, xb, phi, gy[:, ib], gl)
        ginp, cg = conditioner_backward(layer.hyper, inp, gphi.reshape(B, -1))
        gx = np.empty_like(x)
        gx[:, ia] = gy[:, ia] + ginp[:, : len(ia)]
        gx[:, ib] = gxb
        return gx, (None if cc is None else ginp[:, len(ia) :]), LayerGrads(hyper=cg)
    if layer.kind == "elementwise":
        if layer.hyper is None:
            gx, gphi = _uni_backward(layer, x, np.asarray(layer.phi, F), gy, gl)
            return gx, None, LayerGrads(phi=gphi)
        phi = _cond_forward(layer.hyper, cc).reshape(B, D, P)
        gx, gphi = _uni_backward(layer, x, phi, gy, gl)
        gc, cg = conditioner_backward(layer.hyper, cc, gphi.reshape(B, D * P))
        return gx, gc, LayerGrads(hyper=cg)
    if layer.kind == "softclip":
        return softclip_backward(x, gy, np.broadcast_to(gl[:, None], x.shape), layer.bound), None, LayerGrads()
    if layer.kind == "permutation":
        gx = np.zeros_like(x)
        gx[:, np.asarray(layer.order)] = gy  # y[:, j] = x[:, order[j]]
        return gx, None, LayerGrads()
    if layer.kind == "rotation":
        R = np.asarray(layer.R, F)
        return gy @ R, None, LayerGrads(R=gy.T @ x)  # y = x R^T
    raise ValueError(layer.kind)


def flow_backward(spec: O.FlowSpec, x, c=None, g_log_prob=None, g_z=None, g_ladj=None):
    """Reverse mode of ``FlowSpec.log_prob`` (weights g_log_prob (B)) and / or of
    ``FlowSpec.forward`` (g_z (B, D), g_ladj (B)).  Returns (gx, gc or None, [LayerGrads])."""
    x = np.asarray(x, F)
    B, D = x.shape
    zs = [x]
    for layer in spec.layers:
        y, _ = layer.forward(zs[-1], c, F)
        zs.append(np.asarray(y, F))
    g = np.zeros((B, D), F) if g_z is None else np.asarray(g_z, F).copy()
    gl = np.zeros(B, F) if g_ladj is None else np.asarray(g_ladj, F).copy()
    if g_log_prob is not None:
        glp = np.asarray(g_log_prob, F)
        # base.log_prob(z) + ladj (distributions.py:115-119; torch normal.py:87-102); a BoxUniform base
        # has a constant density on its support: no d/dz term
        if spec.base != "uniform":
            g = g + glp[:, None] * (-(zs[-1] - spec.loc) / spec.scale**2)
        gl = gl + glp
    gc = None
    grads = [None] * len(spec.layers)
    for i in reversed(range(len(spec.layers))):
        g, gci, grads[i] = layer_backward(spec.layers[i], zs[i], c, g, gl)
        if gci is not None:
            gc = gci if gc is None else gc + gci
    if gc is not None and c is not None and np.asarray(c).ndim == 1:
        gc = gc.sum(0)
    return g, gc, grads


# --------------------------------------------------------------------------- #
# inverse direction (reparameterised sampling): implicit differentiation
# --------------------------------------------------------------------------- #


def _layer_passes(layer: O.Layer) -> int:
    """Number of Richardson sweeps that solve J^T v = g exactly: J is triangular in the layer's order
    classes, so (I - S J^T) is nilpotent with that index (S = 1 / diag J)."""
    if layer.kind == "autoregressive":
        return max(int(layer.passes), 1)
    if layer.kind == "coupling":
        return 2
    return 1


def layer_inverse_backward(layer: O.Layer, x, c, g):
    """x = layer^{-1}(y; theta, c) at the solution x.  Given g = dL/dx returns
    (v = dL/dy, gc, LayerGrads) by the implicit-function theorem applied to y = f(x; theta, c):
        dL/dy = J^{-T} g,   dL/dtheta = -(df/dtheta)^T v,   dL/dc = -(df/dc)^T v,   J = df/dx.
    (torch.autograd reaches the same numbers by back-propagating through the `passes` sweeps of
    transforms.py:994-1000.)  J^T v is one call of `layer_backward` with gy = v, gl = 0."""
    x = np.asarray(x, F)
    g = np.asarray(g, F)
    B, D = x.shape
    zero = np.zeros(B, F)
    if layer.kind == "permutation":
        v = g[:, np.asarray(layer.order)]  # J = P (orthogonal): J^{-T} g = P g
    elif layer.kind == "rotation":
        v = g @ np.asarray(layer.R, F).T  # J = R (orthogonal): J^{-T} g = R g
    else:
        s = _layer_diag(layer, x, c)  # diag J: derivative of the univariate bijector at fixed parameters
        v = g / s
        for _ in range(_layer_passes(layer) - 1):
            jtv, _, _ = layer_backward(layer, x, c, v, zero)
            v = v + (g - jtv) / s
    _, gc, lg = layer_backward(layer, x, c, -v, zero)
    return v, gc, lg


def _layer_diag(layer: O.Layer, x, c):
    """diag(df/dx): the derivative of the univariate bijector at fixed parameters."""
    B, D = x.shape
    ones, zero = np.ones_like(x), np.zeros_like(x)
    cc = _ctx_rows(c, B)
    P = 3 * layer.bins - 1 if layer.univariate in ("rqs", "crqs") else 2
    if layer.kind == "autoregressive":
        inp = x if cc is None else np.concatenate([x, cc], -1)
        phi = _cond_forward(layer.hyper, inp).reshape(B, D, P)
        s, _ = _uni_backward(layer, x, phi, ones, np.zeros(B, F))
        return s
    if layer.kind == "coupling":
        ia, ib = np.nonzero(layer.mask)[0], np.nonzero(~layer.mask)[0]
        inp = x[:, ia] if cc is None else np.concatenate([x[:, ia], cc], -1)
        phi = _cond_forward(layer.hyper, inp).reshape(B, len(ib), P)
        sb, _ = _uni_backward(layer, x[:, ib], phi, ones[:, ib], np.zeros(B, F))
        s = np.ones_like(x)
        s[:, ib] = sb
        return s
    if layer.kind == "elementwise":
        phi = np.asarray(layer.phi, F) if layer.hyper is None else _cond_forward(layer.hyper, cc).reshape(B, D, P)
        s, _ = _uni_backward(layer, x, phi, ones, np.zeros(B, F))
        return s
    if layer.kind == "softclip":
        return softclip_backward(x, ones, zero, layer.bound)
    raise ValueError(layer.kind)


def flow_inverse_backward(spec: O.FlowSpec, z, c=None, g_x=None, g_log_prob=None):
    """Reverse mode of ``x = FlowSpec.inverse(z)`` (weights g_x (B, D)) and, with ``g_log_prob``, of the
    second output of ``FlowSpec.inverse_and_log_prob`` (distributions.py:129-138).
    Returns (gz, gc or None, [LayerGrads])."""
    z = np.asarray(z, F)
    B, D = z.shape
    x = spec.inverse(z, c, F)
    xs = [np.asarray(x, F)]  # xs[l] = input of layer l in the forward direction
    for layer in spec.layers:
        y, _ = layer.forward(xs[-1], c, F)
        xs.append(np.asarray(y, F))
    g = np.zeros((B, D), F) if g_x is None else np.asarray(g_x, F).copy()
    n = len(spec.layers)
    grads = [LayerGrads() for _ in range(n)]
    gc_tot = None
    gz = np.zeros((B, D), F)

    def add_grads(dst: LayerGrads, src: LayerGrads):
        if src.hyper is not None:
            if dst.hyper is None:
                dst.hyper = CondGrads([np.zeros_like(w) for w in src.hyper.weights],
                                      [None if b is None else np.zeros_like(b) for b in src.hyper.biases])  # fmt: skip
            for i in range(len(src.hyper.weights)):
                dst.hyper.weights[i] += src.hyper.weights[i]
                if src.hyper.biases[i] is not None:
                    dst.hyper.biases[i] += src.hyper.biases[i]
        if src.phi is not None:
            dst.phi = src.phi if dst.phi is None else dst.phi + src.phi
        if src.R is not None:
            dst.R = src.R if dst.R is None else dst.R + src.R

    if g_log_prob is not None:
        # log p = base.log_prob(z) + sum_l ladj_l(x_l; theta): explicit part at fixed x, and d/dx joins g
        glp = np.asarray(g_log_prob, F)
        gx_l, gc_l, lgs = flow_backward(spec, x, c, g_ladj=glp)
        g = g + gx_l
        for i in range(n):
            add_grads(grads[i], lgs[i])
        if gc_l is not None:
            gc_tot = gc_l if np.asarray(c).ndim > 1 else gc_l  # flow_backward already sums a broadcast row
        if spec.base != "uniform":
            gz = gz + glp[:, None] * (-(z - spec.loc) / spec.scale**2)
    gc_rows = None
    for i in range(n):
        g, gci, lg = layer_inverse_backward(spec.layers[i], xs[i], c, g)
        add_grads(grads[i], lg)
        if gci is not None:
            gc_rows = gci if gc_rows is None else gc_rows + gci
    if gc_rows is not None:
        if c is not None and np.asarray(c).ndim == 1:
            gc_rows = gc_rows.sum(0)
        gc_tot = gc_rows if gc_tot is None else gc_tot + gc_rows
    return gz + g, gc_tot, grads
