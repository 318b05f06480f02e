"""Generates the GRADIENT golden vectors tests/golden/grad_*.npz from the UNMODIFIED reference.

Run in the build container only (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_grad.py

For every case it rebuilds the reference flow exactly as tests/golden/make_golden.py did (same
seed, same constructor, same inputs — read back from the flow_*.npz that script wrote), casts it
to fp64 and lets ``torch.autograd`` differentiate

* ``L1 = sum_b g_b * flow(c).log_prob(x)_b``                                (training loss shape)
* ``L2 = sum(gz * z) + sum(gl * ladj)`` with ``z, ladj = flow(c).transform.call_and_ladj(x)``

w.r.t. x, c and every parameter.  Small cases store the full parameter gradients; the BASELINE
configs (millions of parameters) store a strided sample of 2048 entries plus the L2 norm of each.
"""

from __future__ import annotations

import copy
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parent))

from zuko.flows import MAF, NICE, NSF  # noqa: E402

OUT = Path(__file__).resolve().parent
SAMPLE = 2048


def gen(seed: int, *shape: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def sample_idx(n: int) -> np.ndarray:
    if n <= SAMPLE:
        return np.arange(n)
    return (np.arange(SAMPLE, dtype=np.int64) * (n // SAMPLE)) % n


def grads_of(flow64, x, c, loss_fn, full: bool, prefix: str) -> dict:
    x = x.clone().requires_grad_(True)
    cc = None if c is None else c.clone().requires_grad_(True)
    for p in flow64.parameters():
        p.grad = None
    loss = loss_fn(flow64, x, cc)
    loss.backward()
    out = {prefix + "gx": x.grad.numpy().copy()}
    if cc is not None:
        out[prefix + "gc"] = cc.grad.numpy().copy()
    for n, p in flow64.named_parameters():
        g = torch.zeros_like(p) if p.grad is None else p.grad
        g = g.detach().numpy().reshape(-1)
        if full:
            out[f"{prefix}pg/{n}"] = g.copy()
        else:
            out[f"{prefix}pg_sample/{n}"] = g[sample_idx(g.size)].copy()
        out[f"{prefix}pg_norm/{n}"] = np.float64(np.linalg.norm(g))
    return out


def grad_case(name: str, build, rows: int, *, seed=0, full=True, w_scale=1.0):
    src = np.load(OUT / f"flow_{name}.npz")
    if seed is not None:
        torch.manual_seed(seed)
    flow = build().eval()
    if w_scale != 1.0:
        with torch.no_grad():
            for p in flow.parameters():
                p.mul_(w_scale)
    f64 = copy.deepcopy(flow).double()
    x = torch.from_numpy(src["x"][:rows]).double()
    c = None
    if "c" in src.files:
        c = torch.from_numpy(src["c"])
        c = (c if c.dim() == 1 else c[:rows]).double()
    B, D = x.shape
    # sanity: the rebuilt flow is the one make_golden.py evaluated
    with torch.no_grad():
        lp = f64(c).log_prob(x)
    assert np.allclose(lp.numpy(), src["log_prob64"][:rows], rtol=1e-12, atol=1e-12), name
    g = gen(9001, B).double()
    gz = gen(9002, B, D).double()
    gl = gen(9003, B).double()
    info = {"rows": np.int64(rows), "g": g.numpy(), "gz": gz.numpy(), "gl": gl.numpy()}
    def l1(f, xx, cc):
        return (g.to(xx.dtype) * f(cc).log_prob(xx)).sum()

    def l2(f, xx, cc):
        z, ladj = f(cc).transform.call_and_ladj(xx)
        return (gz.to(xx.dtype) * z).sum() + (gl.to(xx.dtype) * ladj).sum()

    info.update(grads_of(f64, x, c, l1, full, "lp/"))
    info.update(grads_of(f64, x, c, l2, full, "tr/"))
    # the reference's OWN fp32 deviation from its fp64 gradients, per tensor, relative to the
    # largest fp64 entry: calibrates the parity bar of the fp32 engine (like log_prob32 / log_prob64
    # do for the forward pass)
    c32 = None if c is None else c.float()
    for prefix, fn in (("lp/", l1), ("tr/", l2)):
        g32 = grads_of(flow, x.float(), c32, fn, True, "")
        g64 = grads_of(f64, x, c, fn, True, "")
        for k, v64 in g64.items():
            if k.startswith("pg_norm/"):
                continue
            scale = max(float(np.abs(v64).max()), 1e-30)
            key = k[3:] if k.startswith("pg/") else k
            info[f"{prefix}err32/{key}"] = np.float64(np.abs(g32[k].astype(np.float64) - v64).max() / scale)
    np.savez_compressed(OUT / f"grad_{name}.npz", **info)
    print(f"grad_{name}: rows={rows} D={D} |gx|max={np.abs(info['lp/gx']).max():.3e}")


def composed_grad_cases():
    """The two user-composed stacks of make_golden.composed_case (rebuilt from the stored tensors)."""
    import zuko
    from functools import partial
    from zuko.flows import ElementWiseTransform, GeneralCouplingTransform
    from zuko.flows.autoregressive import MaskedAutoregressiveTransform
    from zuko.lazy import Flow, UnconditionalDistribution, UnconditionalTransform
    from zuko.transforms import MonotonicRQSTransform, PermutationTransform, RotationTransform, SoftclipTransform

    D, C = 5, 3

    def build_composed():
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

    def build_uncond():
        torch.manual_seed(4)
        layers = [
            ElementWiseTransform(D, 0, univariate=partial(MonotonicRQSTransform, slope=1e-3), shapes=[(4,), (4,), (3,)]),
            MaskedAutoregressiveTransform(D, 0, hidden_features=[32, 32]),
            UnconditionalTransform(SoftclipTransform, bound=11.0),
            ElementWiseTransform(D, 0),
        ]
        base = UnconditionalDistribution(zuko.distributions.DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)
        return Flow(layers, base)

    grad_case("composed", build_composed, 96, seed=None)
    grad_case("composed_uncond", build_uncond, 96, seed=None)


def unit_grad_cases():
    """Element-level gradients of the univariate bijectors on the inputs of units.npz."""
    from zuko.transforms import MonotonicAffineTransform, MonotonicRQSTransform, SoftclipTransform

    u = np.load(OUT / "units.npz")
    out = {}
    K = 8
    for tag in ("s01", "s1", "s3"):
        phi = torch.from_numpy(u[f"rqs_{tag}_phi"]).double().requires_grad_(True)
        x = torch.from_numpy(u[f"rqs_{tag}_x"]).double()
        x[~torch.isfinite(x) | (x.abs() > 1e20)] = 7.0  # keep the loss finite
        x.requires_grad_(True)
        gy, gl = gen(9100, *x.shape).double(), gen(9101, *x.shape).double()
        t = MonotonicRQSTransform(phi[..., :K], phi[..., K : 2 * K], phi[..., 2 * K :])
        y, ladj = t.call_and_ladj(x)
        ((gy * y).sum() + (gl * ladj).sum()).backward()
        out[f"rqs_{tag}_x"] = x.detach().numpy()
        out[f"rqs_{tag}_gy"], out[f"rqs_{tag}_gl"] = gy.numpy(), gl.numpy()
        out[f"rqs_{tag}_gx"], out[f"rqs_{tag}_gphi"] = x.grad.numpy(), phi.grad.numpy()
    for K2 in (16, 5):
        phi = torch.from_numpy(u[f"rqs_shared{K2}_phi"]).double().requires_grad_(True)
        x = torch.from_numpy(u[f"rqs_shared{K2}_x"]).double().requires_grad_(True)
        gy, gl = gen(9200 + K2, *x.shape).double(), gen(9201 + K2, *x.shape).double()
        t = MonotonicRQSTransform(phi[..., :K2], phi[..., K2 : 2 * K2], phi[..., 2 * K2 :])
        y, ladj = t.call_and_ladj(x)
        ((gy * y).sum() + (gl * ladj).sum()).backward()
        out[f"rqs_shared{K2}_gy"], out[f"rqs_shared{K2}_gl"] = gy.numpy(), gl.numpy()
        out[f"rqs_shared{K2}_gx"], out[f"rqs_shared{K2}_gphi"] = x.grad.numpy(), phi.grad.numpy()
    phi = torch.from_numpy(u["affine_phi"]).double().requires_grad_(True)
    x = torch.from_numpy(u["affine_x"]).double().requires_grad_(True)
    gy, gl = gen(9300, *x.shape).double(), gen(9301, *x.shape).double()
    t = MonotonicAffineTransform(phi[..., 0], phi[..., 1])
    y, ladj = t.call_and_ladj(x)
    ((gy * y).sum() + (gl * ladj).sum()).backward()
    out["affine_gy"], out["affine_gl"] = gy.numpy(), gl.numpy()
    out["affine_gx"], out["affine_gphi"] = x.grad.numpy(), phi.grad.numpy()
    x0 = torch.from_numpy(u["softclip_x"]).double()
    gy, gl = gen(9400, *x0.shape).double(), gen(9401, *x0.shape).double()
    out["softclip_gy"], out["softclip_gl"] = gy.numpy(), gl.numpy()
    for b in (1.0, 11.0):
        x = x0.clone().requires_grad_(True)
        t = SoftclipTransform(bound=b)
        y = t(x)
        ladj = t.log_abs_det_jacobian(x, y)
        ((gy * y).sum() + (gl * ladj).sum()).backward()
        out[f"softclip{int(b)}_gx"] = x.grad.numpy()
    np.savez_compressed(OUT / "grad_units.npz", **out)
    print("grad_units: done")


def main():
    torch.set_num_threads(os.cpu_count())
    unit_grad_cases()
    grad_case("cfg1_maf", lambda: MAF(4, 0, transforms=2, hidden_features=[32, 32]), 256)
    grad_case("cfg2_nsf", lambda: NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 256, full=False)
    grad_case("cfg3_maf", lambda: MAF(32, 0, transforms=8, hidden_features=[512] * 4), 64, full=False)
    grad_case("cfg4_nsf", lambda: NSF(64, 0, transforms=8, bins=16), 32, full=False)
    grad_case("cfg5_nsf", lambda: NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3), 32, full=False)
    grad_case("nsf35_row", lambda: NSF(3, 5), 128)
    grad_case("maf35_batch", lambda: MAF(3, 5), 128)
    grad_case("nice35", lambda: NICE(3, 5), 128)
    grad_case("nsf5_passes2", lambda: NSF(5, 0, passes=2, hidden_features=[32, 32]), 128)
    grad_case("maf5_randperm", lambda: MAF(5, 2, randperm=True, hidden_features=[24]), 128, seed=7)
    grad_case("nsf1_elementwise", lambda: NSF(1, 3, hidden_features=[16]), 128)
    grad_case("nsf6_stress", lambda: NSF(6, 3, transforms=3, hidden_features=[64, 64]), 256, w_scale=3.0)
    composed_grad_cases()


if __name__ == "__main__":
    main()
