"""Shared test helpers: golden loading and construction of the zuko_b200 counterparts of
the flows that tests/golden/make_golden.py built with the reference."""

from __future__ import annotations

import zlib
from functools import partial
from pathlib import Path

import numpy as np
import torch

import zuko_b200 as zuko
from zuko_b200.flows import MAF, NCSF, NICE, NSF, ElementWiseTransform, GeneralCouplingTransform, MaskedAutoregressiveTransform
from zuko_b200.lazy import Flow, UnconditionalDistribution, UnconditionalTransform
from zuko_b200.transforms import MonotonicRQSTransform, PermutationTransform, RotationTransform, SoftclipTransform

GOLDEN = Path(__file__).resolve().parent / "golden"

# name -> (seed, constructor) — must mirror tests/golden/make_golden.py:main()
FLOW_CASES = {
    "cfg1_maf": (0, lambda: MAF(4, 0, transforms=2, hidden_features=[32, 32])),
    "cfg2_nsf": (0, lambda: NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3)),
    "cfg3_maf": (0, lambda: MAF(32, 0, transforms=8, hidden_features=[512] * 4)),
    "cfg4_nsf": (0, lambda: NSF(64, 0, transforms=8, bins=16)),
    "cfg5_nsf": (0, lambda: NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3)),
    "nsf35_row": (0, lambda: NSF(3, 5)),
    "maf35_batch": (0, lambda: MAF(3, 5)),
    "nice35": (0, lambda: NICE(3, 5)),
    "nsf5_passes2": (0, lambda: NSF(5, 0, passes=2, hidden_features=[32, 32])),
    "maf5_randperm": (7, lambda: MAF(5, 2, randperm=True, hidden_features=[24])),
    "nsf1_elementwise": (0, lambda: NSF(1, 3, hidden_features=[16])),
    "nsf6_stress": (0, lambda: NSF(6, 3, transforms=3, hidden_features=[64, 64])),
    "ncsf34": (0, lambda: NCSF(3, 4, hidden_features=[32, 32])),
    # non-ReLU activations (tests/golden/make_golden_act.py)
    "act_maf_elu": (0, lambda: MAF(5, 2, transforms=2, hidden_features=[32, 32], activation=torch.nn.ELU)),
    "act_nsf_tanh": (0, lambda: NSF(4, 0, transforms=2, hidden_features=[64, 64], activation=torch.nn.Tanh)),
    "act_nsf_silu": (0, lambda: NSF(3, 2, transforms=2, hidden_features=[32], activation=torch.nn.SiLU)),
    "act_maf_gelu": (0, lambda: MAF(4, 0, transforms=2, hidden_features=[64, 64], activation=torch.nn.GELU)),
    "act_nice_lrelu": (0, lambda: NICE(4, 2, hidden_features=[32], activation=torch.nn.LeakyReLU)),
    "act_maf_softplus": (0, lambda: MAF(3, 0, transforms=2, hidden_features=[16], activation=torch.nn.Softplus)),
    "act_maf_sigmoid": (0, lambda: MAF(3, 1, transforms=2, hidden_features=[16, 16], activation=torch.nn.Sigmoid)),
    # residual conditioners (tests/golden/make_golden_res.py)
    "res_nsf_relu": (0, lambda: NSF(5, 2, transforms=2, residual=True, hidden_features=[32, 32])),
    "res_maf_elu": (0, lambda: MAF(4, 0, transforms=2, residual=True, hidden_features=[24], activation=torch.nn.ELU)),
    "res_nsf_mixed": (0, lambda: NSF(3, 1, transforms=2, residual=True, hidden_features=[16, 32, 32])),
}
ACT_CASES = [k for k in FLOW_CASES if k.startswith(("act_", "res_"))]
SMALL_CASES = [k for k in FLOW_CASES if not k.startswith(("cfg2", "cfg3", "cfg4", "cfg5"))]
BIG_CASES = ["cfg2_nsf", "cfg3_maf", "cfg4_nsf", "cfg5_nsf"]


def load(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz", allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def crc(t: torch.Tensor) -> int:
    return zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())


def _composed():
    D, C = 5, 3
    torch.manual_seed(3)
    order = torch.randperm(D)
    A = torch.randn(D, D)
    layers = [
        MaskedAutoregressiveTransform(D, C, hidden_features=[32, 32]),
        UnconditionalTransform(SoftclipTransform, bound=11.0),
        UnconditionalTransform(PermutationTransform, order, buffer=True),
        MaskedAutoregressiveTransform(D, C, univariate=partial(MonotonicRQSTransform, slope=1e-3),
                                      shapes=[(8,), (8,), (7,)], hidden_features=[32, 32]),  # fmt: skip
        UnconditionalTransform(RotationTransform, A, buffer=True),
        GeneralCouplingTransform(D, C, hidden_features=[32]),
        ElementWiseTransform(D, C, hidden_features=[16]),
    ]
    base = UnconditionalDistribution(zuko.distributions.DiagNormal, torch.zeros(D) + 0.25, torch.ones(D) * 1.5, buffer=True)
    return Flow(layers, base)


def _composed_uncond():
    D = 5
    torch.manual_seed(4)
    layers = [
        ElementWiseTransform(D, 0, univariate=partial(MonotonicRQSTransform, slope=1e-3), shapes=[(4,), (4,), (3,)]),
        MaskedAutoregressiveTransform(D, 0, hidden_features=[32, 32]),
        UnconditionalTransform(SoftclipTransform, bound=11.0),
        ElementWiseTransform(D, 0),
    ]
    base = UnconditionalDistribution(zuko.distributions.DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)
    return Flow(layers, base)


FLOW_CASES["composed"] = (None, _composed)
FLOW_CASES["composed_uncond"] = (None, _composed_uncond)
SMALL_CASES += ["composed", "composed_uncond"]


def build_flow(name: str, golden: dict | None = None):
    """Builds the zuko_b200 flow for a golden case under the same seed as the reference,
    then (when the case stores them) loads the reference's own tensors, and applies the
    case's weight scaling.  Returns the flow in eval mode on the CPU."""
    golden = golden if golden is not None else load(f"flow_{name}")
    seed, ctor = FLOW_CASES[name]
    if seed is not None:
        torch.manual_seed(seed)
    flow = ctor().eval()
    stored = {k[3:]: torch.from_numpy(v) for k, v in golden.items() if k.startswith("sd/")}
    if stored:
        flow.load_state_dict(stored, strict=True)
    if "w_scale" in golden:
        with torch.no_grad():
            for p in flow.parameters():
                p.mul_(float(golden["w_scale"]))
    return flow


def rel_err(ours: np.ndarray, ref: np.ndarray) -> float:
    ours, ref = np.asarray(ours, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(ours - ref) / np.maximum(np.abs(ref), 1.0))) if ours.size else 0.0


def assert_log_prob_parity(ours, g: dict, rtol: float = 1e-5):
    """The parity bar of BASELINE.json / SURVEY §7.4-1b:
    |ours - ref64| <= max(rtol * |ref64|, 2 * |ref32 - ref64|) per sample (the second term
    only matters for sharp splines where the reference's own fp32 answer drifts), with an
    absolute floor rtol * 1 for log-densities that happen to be ~0."""
    ours = np.asarray(ours, dtype=np.float64)
    ref64, ref32 = g["log_prob64"].astype(np.float64), g["log_prob32"].astype(np.float64)
    tol = np.maximum(rtol * np.maximum(np.abs(ref64), 1.0), 2.0 * np.abs(ref32 - ref64))
    err = np.abs(ours - ref64)
    if "w_scale" in g:
        # stress set (weights x3 => very sharp splines): no two fp32 implementations agree per
        # sample.  Require instead that the engine is no less accurate than the
        # reference's own fp32 path in distribution: max, 99th percentile and median.
        ref_err = np.abs(ref32 - ref64)
        for q in (100, 99, 50):
            ours_q, ref_q = np.percentile(err, q), np.percentile(ref_err, q)
            assert ours_q <= max(2.0 * ref_q, rtol * np.percentile(np.abs(ref64), q)), (q, ours_q, ref_q)
        return
    worst = int(np.argmax(err - tol))
    assert np.all(err <= tol), (
        f"log_prob parity: sample {worst}: ours={ours[worst]!r} ref64={ref64[worst]!r} ref32={ref32[worst]!r} "
        f"err={err[worst]:.3e} tol={tol[worst]:.3e}"
    )


# --------------------------------------------------------------------------- #
# gradients
# --------------------------------------------------------------------------- #

GRAD_CASES_FULL = ["cfg1_maf", "nsf35_row", "maf35_batch", "nice35", "nsf5_passes2", "maf5_randperm",
                   "nsf1_elementwise", "nsf6_stress", "composed", "composed_uncond", "ncsf34", *ACT_CASES]  # fmt: skip
GRAD_CASES_SAMPLED = ["cfg2_nsf", "cfg3_maf", "cfg4_nsf", "cfg5_nsf"]
GRAD_SAMPLE = 2048


def grad_sample_idx(n: int) -> np.ndarray:
    """Mirror of tests/golden/make_golden_grad.py:sample_idx."""
    if n <= GRAD_SAMPLE:
        return np.arange(n)
    return (np.arange(GRAD_SAMPLE, dtype=np.int64) * (n // GRAD_SAMPLE)) % n


def grad_inputs(name: str):
    """(golden-grad dict, x, c) of a gradient case: the first `rows` rows of the forward case."""
    gg = load(f"grad_{name}")
    g = load(f"flow_{name}")
    rows = int(gg["rows"])
    x = g["x"][:rows]
    c = g.get("c")
    if c is not None and c.ndim == 2:
        c = c[:rows]
    return gg, x, c


def oracle_named_grads(flow, layer_grads) -> dict:
    """Maps oracle_grad.LayerGrads (one per member of the composed transform) onto the
    state-dict parameter names of the flow module."""
    out = {}
    for i, (t, lg) in enumerate(zip(flow.transform.transforms, layer_grads, strict=True)):
        pre = f"transform.transforms.{i}."
        if lg.hyper is not None:
            lins = []  # (state-dict index, linear module) in execution order, residual blocks flattened
            for j, m in enumerate(t.hyper):
                if type(m).__name__ == "Residual":
                    lins += [(f"{j}.{k}", inner) for k, inner in enumerate(m) if hasattr(inner, "weight")]
                elif hasattr(m, "weight"):
                    lins.append((j, m))
            for (j, m), gw, gb in zip(lins, lg.hyper.weights, lg.hyper.biases, strict=True):
                out[f"{pre}hyper.{j}.weight"] = np.asarray(gw).reshape(-1)
                if m.bias is not None:
                    out[f"{pre}hyper.{j}.bias"] = np.asarray(gb).reshape(-1)
        if lg.phi is not None:
            col = 0
            D = lg.phi.shape[0]
            for k, p in enumerate(t.phi):
                w = int(np.prod(p.shape[1:])) if p.dim() > 1 else 1
                out[f"{pre}phi.{k}"] = np.asarray(lg.phi[:, col : col + w]).reshape(-1)
                col += w
            assert col == lg.phi.shape[1] and D == t.features
    return out


def grad_bar(gg: dict, prefix: str, name: str, rtol: float, ref32_factor: float) -> float:
    """Parity bar of one gradient tensor: ``max(rtol, ref32_factor * e32)`` where e32 is the
    reference's OWN fp32-vs-fp64 deviation on that tensor (stored by make_golden_grad.py), the
    same calibration the forward parity rule uses (assert_log_prob_parity)."""
    key = f"{prefix}err32/{name}"
    return max(rtol, ref32_factor * float(gg[key])) if (ref32_factor and key in gg) else rtol


def assert_param_grads(named: dict, gg: dict, prefix: str, rtol: float, what: str = "", ref32_factor: float = 0.0,
                       minus: dict | None = None):  # fmt: skip
    """Compares a name -> flat-gradient dict with the golden gradients (full or sampled).
    Error is measured relative to the largest entry of each tensor's golden gradient
    (`|ours - ref| <= bar * max|ref|`), the natural scale of a summed-over-batch gradient."""
    checked = 0
    for key in gg:
        if key.startswith(prefix + "pg/"):
            name, ref, full = key[len(prefix) + 3 :], gg[key], True
        elif key.startswith(prefix + "pg_sample/"):
            name, ref, full = key[len(prefix) + 10 :], gg[key], False
        else:
            continue
        assert name in named, f"{what}: no gradient produced for {name}"
        ours = np.asarray(named[name], dtype=np.float64).reshape(-1)
        if minus is not None and name in minus:  # golden sum minus the rows that were given zero weight
            corr = np.asarray(minus[name], dtype=np.float64).reshape(-1)
            ref = ref - (corr if full else corr[grad_sample_idx(corr.size)])
        if not full:
            ours = ours[grad_sample_idx(ours.size)]
        scale = max(float(np.abs(ref).max()), 1e-30)
        err = float(np.abs(ours - ref).max()) / scale
        bar = grad_bar(gg, prefix, name, rtol, ref32_factor)
        assert err <= bar, f"{what}: d/d{name}: max err {err:.3e} of max|grad| {scale:.3e} (bar {bar:.1e})"
        checked += 1
    assert checked > 0, f"{what}: golden file holds no parameter gradients under {prefix}"
    return checked


def relu_kink_rows(spec, x, c, tau: float = 1e-5, tau_knot: float = 5e-5) -> np.ndarray:
    """Rows of the batch on which the flow is NOT differentiable to working precision (fp64 oracle):

    * ReLU kinks — some hidden unit of some conditioner has a pre-activation
      ``|p| < tau * (sum_k |a_k w_k| + |b|)``, i.e. within the rounding error of the split-bf16 GEMM
      of zero (measured on the BASELINE configs by emulating the 3-term bf16 product: median 7e-7,
      99.99th percentile 1e-5 of that sum; every sign flip observed on the GPU sat below 5e-7);
    * spline knots — some input of a rational-quadratic spline lies within ``tau_knot`` of one of its
      knots: the spline is C1, so d(ladj)/dx and d(ladj)/d(phi) jump across a knot (the engine's
      layer inputs carry ~1e-5 of forward error, the knots ~1e-6).

    There the gradient jumps by a finite amount, so two correct implementations whose intermediates
    differ in the last bits (the reference's own fp32 path included) may return either one-sided
    gradient.  The gradient parity tests give those rows zero weight and correct the golden sums
    with the (pinned) gradient oracle — gradients are linear in the per-row weights."""
    x = np.asarray(x, np.float64)
    B = x.shape[0]
    bad = np.zeros(B, dtype=bool)
    z = x
    for layer in spec.layers:
        cond = layer.hyper
        phi = None
        zt = z  # inputs of the univariate bijector
        if cond is not None:
            cc = None if c is None else (np.broadcast_to(c, (B, np.asarray(c).shape[-1])) if np.asarray(c).ndim == 1 else np.asarray(c, np.float64))
            if layer.kind == "autoregressive":
                h = z if cc is None else np.concatenate([z, cc], -1)
            elif layer.kind == "coupling":
                za = z[:, np.nonzero(layer.mask)[0]]
                zt = z[:, np.nonzero(~layer.mask)[0]]
                h = za if cc is None else np.concatenate([za, cc], -1)
            else:
                h = cc
            acts_, _res = cond.flags()
            a_, outs_ = cond.trace(h)
            phi = outs_[-1]
            for i, name_ in enumerate(acts_):
                if name_ in ("ReLU", "LeakyReLU"):  # the only supported activations with a kink
                    W = cond.weights[i] * (1.0 if cond.masks[i] is None else cond.masks[i])
                    bias = 0.0 if cond.biases[i] is None else cond.biases[i]
                    scale = np.abs(a_[i]) @ np.abs(W).T + np.abs(bias) + (np.abs(a_[i - 1]) if _res[i] else 0.0)
                    bad |= (np.abs(outs_[i]) < tau * scale).any(-1)
        elif layer.kind == "elementwise":
            phi = np.broadcast_to(np.asarray(layer.phi, np.float64), (B, *np.asarray(layer.phi).shape))
        if phi is not None and layer.univariate in ("rqs", "crqs"):
            P = 3 * layer.bins - 1
            X, _, _ = O_rqs_knots(phi.reshape(B, zt.shape[1], P), layer.bins, layer.bound, layer.slope)
            if layer.univariate == "crqs":  # the spline sees the circularly shifted input; the shift itself
                from oracle import oracle as _O  # jumps where remainder wraps (x = 2 k bound)

                bad |= (np.abs(np.remainder(zt, 2 * layer.bound)) < tau_knot).any(-1)
                bad |= (np.abs(np.remainder(zt, 2 * layer.bound) - 2 * layer.bound) < tau_knot).any(-1)
                zt = _O.circular_shift(zt, layer.bound)
            bad |= (np.abs(zt[..., None] - X).min(-1) < tau_knot).any(-1)
        z, _ = layer.forward(z, c, np.float64)
    return bad


def O_act(name):
    from oracle import oracle as O

    return O.ACTIVATIONS[name][0]


def O_rqs_knots(phi, bins, bound, slope):
    from oracle import oracle as O

    return O.rqs_knots(phi, bins, bound, slope)


INV_GRAD_CASES = ["cfg1_maf", "nsf35_row", "maf35_batch", "nice35", "nsf5_passes2", "maf5_randperm", "nsf1_elementwise",
                  "ncsf34", "act_maf_elu", "act_nsf_tanh", "composed", "composed_uncond", "cfg2_nsf",
                  "res_nsf_relu", "res_maf_elu", "res_nsf_mixed"]  # fmt: skip


def inv_grad_inputs(name: str):
    """(golden dict, z, c) of an inverse-direction gradient case (tests/golden/make_golden_inv_grad.py)."""
    gg = load(f"invgrad_{name}")
    c = load(f"flow_{name}").get("c")
    if c is not None and c.ndim == 2:
        c = c[: int(gg["rows"])]
    return gg, gg["z"].astype(np.float32), c
