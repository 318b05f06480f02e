"""Generates the golden vectors under tests/golden/ from the UNMODIFIED Python reference.

Run in the build container only (the reference is not present on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports probabilists/zuko from /root/reference (read-only), builds seeded models with
the reference constructors, evaluates them on seeded inputs in fp32 and fp64
(``copy.deepcopy(flow).double()``), and stores inputs + outputs as small ``.npz`` files.
The oracle (oracle/) and the CUDA engine (zuko_b200/) are both checked against these files
by tests/; nothing here is imported at test time.

Every case stores CRC32 checksums of the reference's parameters and buffers so that the
tests can verify that zuko_b200's own constructors reproduce the reference initialisation
bit for bit under the same ``torch.manual_seed``; small cases also store the tensors.
"""

from __future__ import annotations

import copy
import os
import sys
import zlib
from functools import partial
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import zuko  # noqa: E402  (the reference)
from zuko.flows import MAF, NICE, NSF, ElementWiseTransform, GeneralCouplingTransform  # noqa: E402
from zuko.flows.autoregressive import MaskedAutoregressiveTransform  # noqa: E402
from zuko.lazy import Flow, UnconditionalDistribution, UnconditionalTransform  # noqa: E402
from zuko.nn import MaskedMLP  # noqa: E402
from zuko.transforms import (  # noqa: E402
    MonotonicAffineTransform,
    MonotonicRQSTransform,
    PermutationTransform,
    RotationTransform,
    SoftclipTransform,
)

OUT = Path(__file__).resolve().parent
assert zuko.__version__ == "1.6.0", zuko.__version__


def crc(t: torch.Tensor) -> int:
    return zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())


def state_info(module: torch.nn.Module, store: bool) -> dict:
    out = {}
    sd = module.state_dict()
    out["sd_keys"] = np.array(list(sd.keys()))
    out["sd_crc"] = np.array([crc(v) for v in sd.values()], dtype=np.int64)
    if store:
        for k, v in sd.items():
            out["sd/" + k] = v.detach().cpu().numpy().copy()
    return out


def gen(seed: int, *shape: int, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def flow_case(name, build, B, ctx, *, seed=0, store=False, x_scale=1.0, w_scale=1.0, inverse_rows=0):
    """build(): reference flow; ctx: None | 'row' (C,) | 'batch' (B, C)."""
    torch.manual_seed(seed)
    flow = build().eval()
    info = state_info(flow, store)
    if w_scale != 1.0:
        with torch.no_grad():
            for n, p in flow.named_parameters():
                p.mul_(w_scale)
        info["w_scale"] = np.float64(w_scale)
    D = (flow.base.loc if hasattr(flow.base, "loc") else flow.base.lower).shape[0]  # NCSF: BoxUniform base
    x = gen(1234, B, D, scale=x_scale)
    c = None
    t0 = flow.transform.transforms[0]
    C = getattr(t0, "hyper", [None])[0].weight.shape[1] - D if isinstance(t0, MaskedAutoregressiveTransform) else None
    if ctx is not None:
        C = ctx[1]
        c = gen(4321, C) if ctx[0] == "row" else gen(4321, B, C)
    f64 = copy.deepcopy(flow).double()
    with torch.no_grad():
        lp32 = flow(c).log_prob(x)
        z32, ladj32 = flow(c).transform.call_and_ladj(x)
        c64 = None if c is None else c.double()
        lp64 = f64(c64).log_prob(x.double())
        z64, ladj64 = f64(c64).transform.call_and_ladj(x.double())
        info.update(x=x.numpy(), log_prob32=lp32.numpy(), log_prob64=lp64.numpy(), z32=z32.numpy(),
                    z64=z64.numpy(), ladj32=ladj32.numpy(), ladj64=ladj64.numpy())  # fmt: skip
        if c is not None:
            info["c"] = c.numpy()
        if inverse_rows:
            zin = gen(777, inverse_rows, D)
            ci = None if c is None else (c if c.dim() == 1 else c[:inverse_rows])
            ci64 = None if ci is None else ci.double()
            xi32 = flow(ci).transform.inv(zin)
            xi64 = f64(ci64).transform.inv(zin.double())
            info.update(zin=zin.numpy(), xinv32=xi32.numpy(), xinv64=xi64.numpy())
    np.savez_compressed(OUT / f"flow_{name}.npz", **info)
    rel = ((lp32.double() - lp64).abs() / lp64.abs()).max().item()
    print(f"flow_{name}: B={B} D={D} ref32-vs-ref64 max rel {rel:.2e}")


def composed_case():
    """User-composed Flow([...]) mixing autoregressive layers with softclip, permutation
    and rotation, as README.md:57-75 / flows/neural.py do."""
    torch.manual_seed(3)
    D, C = 5, 3
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
    flow = Flow(layers, base).eval()
    info = state_info(flow, True)
    x, c = gen(11, 96, D), gen(12, 96, C)
    f64 = copy.deepcopy(flow).double()
    with torch.no_grad():
        lp32 = flow(c).log_prob(x)
        lp64 = f64(c.double()).log_prob(x.double())
        z64, ladj64 = f64(c.double()).transform.call_and_ladj(x.double())
        zin = gen(13, 96, D, scale=0.7)
        xi64 = f64(c.double()).transform.inv(zin.double())
    info.update(x=x.numpy(), c=c.numpy(), log_prob32=lp32.numpy(), log_prob64=lp64.numpy(), z64=z64.numpy(),
                ladj64=ladj64.numpy(), zin=zin.numpy(), xinv64=xi64.numpy(), order=order.numpy(), A=A.numpy())  # fmt: skip
    np.savez_compressed(OUT / "flow_composed.npz", **info)
    print("flow_composed: done")

    # unconditional stack with shared per-dimension tables (gaussianization.py:74-77)
    torch.manual_seed(4)
    layers = [
        ElementWiseTransform(D, 0, univariate=partial(MonotonicRQSTransform, slope=1e-3), shapes=[(4,), (4,), (3,)]),
        MaskedAutoregressiveTransform(D, 0, hidden_features=[32, 32]),
        UnconditionalTransform(SoftclipTransform, bound=11.0),
        ElementWiseTransform(D, 0),
    ]
    base = UnconditionalDistribution(zuko.distributions.DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)
    flow = Flow(layers, base).eval()
    info = state_info(flow, True)
    x = gen(21, 96, D)
    f64 = copy.deepcopy(flow).double()
    with torch.no_grad():
        lp32 = flow().log_prob(x)
        lp64 = f64().log_prob(x.double())
        z64, ladj64 = f64().transform.call_and_ladj(x.double())
        zin = gen(23, 96, D, scale=0.7)
        xi64 = f64().transform.inv(zin.double())
    info.update(x=x.numpy(), log_prob32=lp32.numpy(), log_prob64=lp64.numpy(), z64=z64.numpy(),
                ladj64=ladj64.numpy(), zin=zin.numpy(), xinv64=xi64.numpy())  # fmt: skip
    np.savez_compressed(OUT / "flow_composed_uncond.npz", **info)
    print("flow_composed_uncond: done")


def unit_cases():
    out = {}
    N, D, K = 64, 5, 8
    edge = torch.tensor([-5.0, 5.0, -5.0000005, 4.9999995, 1e30, -1e30, 0.0, 4.999, -4.999, 7.5])
    for tag, s in (("s01", 0.1), ("s1", 1.0), ("s3", 3.0)):
        phi = gen(100 + int(s * 10), N, D, 3 * K - 1, scale=s)
        x = gen(200 + int(s * 10), N, D, scale=2.5)
        x[: edge.numel(), 0] = edge
        for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
            p = phi.to(dt)
            t = MonotonicRQSTransform(p[..., :K], p[..., K : 2 * K], p[..., 2 * K :])
            y, ladj = t.call_and_ladj(x.to(dt))
            xi = t.inv(y)
            out[f"rqs_{tag}_y{sfx}"] = y.numpy()
            out[f"rqs_{tag}_ladj{sfx}"] = ladj.numpy()
            out[f"rqs_{tag}_xinv{sfx}"] = xi.numpy()
            # inverse evaluated on fresh points of the codomain as well
            yq = gen(300 + int(s * 10), N, D, scale=2.5).to(dt)
            out[f"rqs_{tag}_inv_of_yq{sfx}"] = t.inv(yq).numpy()
            if sfx == "64":
                out[f"rqs_{tag}_yq"] = yq.float().numpy()
                out[f"rqs_{tag}_horizontal"] = t.horizontal.numpy()
                out[f"rqs_{tag}_vertical"] = t.vertical.numpy()
                out[f"rqs_{tag}_derivatives"] = t.derivatives.numpy()
        out[f"rqs_{tag}_phi"] = phi.numpy()
        out[f"rqs_{tag}_x"] = x.numpy()
    # shared (unbatched) spline parameters, K = 16 and an odd K = 5
    for K2 in (16, 5):
        phi = gen(400 + K2, D, 3 * K2 - 1)
        x = gen(401 + K2, N, D, scale=2.0)
        t = MonotonicRQSTransform(phi[..., :K2].double(), phi[..., K2 : 2 * K2].double(), phi[..., 2 * K2 :].double())
        y, ladj = t.call_and_ladj(x.double())
        out[f"rqs_shared{K2}_phi"] = phi.numpy()
        out[f"rqs_shared{K2}_x"] = x.numpy()
        out[f"rqs_shared{K2}_y64"] = y.numpy()
        out[f"rqs_shared{K2}_ladj64"] = ladj.numpy()
        out[f"rqs_shared{K2}_xinv64"] = t.inv(x.double()).numpy()
    # affine
    phi = gen(500, N, D, 2, scale=2.0)
    x = gen(501, N, D, scale=2.0)
    for dt, sfx in ((torch.float32, "32"), (torch.float64, "64")):
        t = MonotonicAffineTransform(phi[..., 0].to(dt), phi[..., 1].to(dt))
        y, ladj = t.call_and_ladj(x.to(dt))
        out[f"affine_y{sfx}"] = y.numpy()
        out[f"affine_ladj{sfx}"] = ladj.numpy()
        out[f"affine_xinv{sfx}"] = t.inv(x.to(dt)).numpy()
    out["affine_phi"], out["affine_x"] = phi.numpy(), x.numpy()
    # softclip
    x = gen(600, N, D, scale=4.0)
    for b in (1.0, 11.0):
        t = SoftclipTransform(bound=b)
        y = t(x.double())
        out[f"softclip{int(b)}_y64"] = y.numpy()
        out[f"softclip{int(b)}_ladj64"] = t.log_abs_det_jacobian(x.double(), y).numpy()
        out[f"softclip{int(b)}_xinv64"] = t.inv(y).numpy()
    out["softclip_x"] = x.numpy()
    # permutation (bit-exact) and rotation
    g = torch.Generator().manual_seed(700)
    order = torch.randperm(D, generator=g)
    x = gen(701, N, D)
    t = PermutationTransform(order)
    out["perm_order"], out["perm_x"] = order.numpy(), x.numpy()
    out["perm_y"] = t(x).numpy()
    out["perm_xinv"] = t.inv(x).numpy()
    A = gen(702, D, D)
    t = RotationTransform(A.double())
    out["rot_A"], out["rot_R64"] = A.numpy(), t.R.numpy()
    out["rot_y64"] = t(x.double()).numpy()
    out["rot_xinv64"] = t.inv(x.double()).numpy()
    # DiagNormal log-prob
    z = gen(800, N, D, scale=1.5)
    loc, scale = gen(801, D), gen(802, D).abs() + 0.5
    out["dn_z"], out["dn_loc"], out["dn_scale"] = z.numpy(), loc.numpy(), scale.numpy()
    out["dn_lp64"] = zuko.distributions.DiagNormal(loc.double(), scale.double()).log_prob(z.double()).numpy()
    np.savez_compressed(OUT / "units.npz", **out)
    print("units: done")


def mask_cases():
    """Bit-exact mask / order construction (zuko/nn.py:258-318, flows/autoregressive.py:106-152)."""
    out = {}
    specs = {
        "maf4": dict(features=4, context=0, hidden_features=[32, 32]),
        "nsf16c8": dict(features=16, context=8, hidden_features=[256] * 3, univariate=MonotonicRQSTransform, shapes=[(8,), (8,), (7,)]),
        "passes2": dict(features=5, context=7, passes=2, hidden_features=[16, 24]),
        "order": dict(features=5, context=0, order=[3, 0, 4, 1, 2], hidden_features=[17]),
        "rev64": dict(features=64, context=0, order=list(range(63, -1, -1)), hidden_features=[64, 64], univariate=MonotonicRQSTransform, shapes=[(16,), (16,), (15,)]),
    }
    g = torch.Generator().manual_seed(5)
    adjacency = torch.rand((5, 5), generator=g) < 0.25
    adjacency = adjacency + torch.eye(5, dtype=bool)
    adjacency = torch.tril(adjacency)
    adjacency[1, 0] = True
    perm = torch.randperm(5, generator=g)
    adjacency = adjacency[perm, :][:, perm]
    specs["adjacency"] = dict(features=5, context=0, adjacency=adjacency, hidden_features=[12, 12])
    adj_ctx = torch.cat((adjacency, torch.rand((5, 3), generator=g) < 0.5), dim=1)
    specs["adjacency_ctx"] = dict(features=5, context=3, adjacency=adj_ctx, hidden_features=[12])
    out["adjacency"], out["adjacency_ctx"] = adjacency.numpy(), adj_ctx.numpy()
    for name, kw in specs.items():
        torch.manual_seed(0)
        t = MaskedAutoregressiveTransform(**kw)
        out[f"{name}/passes"] = np.int64(t.passes)
        if t.order is not None:
            out[f"{name}/order"] = t.order.numpy()
        for i, m in enumerate(l for l in t.hyper if hasattr(l, "mask")):
            out[f"{name}/mask{i}"] = np.packbits(m.mask.numpy(), axis=None)
            out[f"{name}/shape{i}"] = np.array(m.mask.shape)
    # a free-standing MaskedMLP on a random adjacency (tests/test_nn.py:39-60)
    adj = torch.rand((4, 3), generator=g) < 0.5
    adj[0, 0] = True
    net = MaskedMLP(adj, [16, 32])
    out["free/adjacency"] = adj.numpy()
    for i, m in enumerate(l for l in net if hasattr(l, "mask")):
        out[f"free/mask{i}"] = m.mask.numpy()
    np.savez_compressed(OUT / "masks.npz", **out)
    print("masks: done")


def main():
    torch.set_num_threads(os.cpu_count())
    unit_cases()
    mask_cases()
    # BASELINE.json configs (reduced batch; cfg1 at its full batch)
    flow_case("cfg1_maf", lambda: MAF(4, 0, transforms=2, hidden_features=[32, 32]), 1024, None, store=True, inverse_rows=64)
    flow_case("cfg2_nsf", lambda: NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 256, ("batch", 8), inverse_rows=16)
    flow_case("cfg3_maf", lambda: MAF(32, 0, transforms=8, hidden_features=[512] * 4), 128, None, inverse_rows=8)
    flow_case("cfg4_nsf", lambda: NSF(64, 0, transforms=8, bins=16), 32, None, inverse_rows=32)
    flow_case("cfg5_nsf", lambda: NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3), 64, ("batch", 16))
    # reference test-suite shapes (tests/test_flows.py:13-94: F(3, 5), broadcast context (5,))
    flow_case("nsf35_row", lambda: NSF(3, 5), 256, ("row", 5), store=True, inverse_rows=64)
    flow_case("maf35_batch", lambda: MAF(3, 5), 256, ("batch", 5), store=True, inverse_rows=64)
    flow_case("nice35", lambda: NICE(3, 5), 256, ("batch", 5), store=True, inverse_rows=64)
    flow_case("nsf5_passes2", lambda: NSF(5, 0, passes=2, hidden_features=[32, 32]), 128, None, store=True, inverse_rows=32)
    flow_case("maf5_randperm", lambda: MAF(5, 2, randperm=True, hidden_features=[24]), 128, ("batch", 2), seed=7, store=True, inverse_rows=32)
    flow_case("nsf1_elementwise", lambda: NSF(1, 3, hidden_features=[16]), 128, ("batch", 3), store=True, inverse_rows=32)
    # stress: sharper splines (weights x3) and inputs beyond the spline domain (x x3)
    flow_case("nsf6_stress", lambda: NSF(6, 3, transforms=3, hidden_features=[64, 64]), 512, ("batch", 3), store=True,
              x_scale=3.0, w_scale=3.0, inverse_rows=64)  # fmt: skip
    composed_case()


if __name__ == "__main__":
    main()
