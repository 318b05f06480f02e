"""Diagnostic: error of the inverse-direction gradients per case (no assertions)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import INV_GRAD_CASES, build_flow, grad_sample_idx, inv_grad_inputs, relu_kink_rows  # noqa: E402
from oracle import oracle as O  # noqa: E402

dev = torch.device("cuda:0")
for name in INV_GRAD_CASES:
    gg, z, c = inv_grad_inputs(name)
    spec = O.flowspec_from_module(build_flow(name))
    kink = relu_kink_rows(spec, gg["x64"], c)
    flow = build_flow(name).to(dev)
    for mode in ("inv", "invlp"):
        for p in flow.parameters():
            p.grad = None
        zt = torch.from_numpy(z).to(dev).requires_grad_()
        ct = None if c is None else torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32)).to(dev).requires_grad_()
        d = flow(ct)
        w = torch.from_numpy(gg["w"]).float().to(dev)
        if mode == "inv":
            x = d.transform.inv(zt)
            loss = (w * x).sum()
        else:
            call, ctx = d._flow_call()
            x, lp = call.inverse(zt, ctx, with_log_prob=True)
            loss = (w * x).sum() + (torch.from_numpy(gg["wl"]).float().to(dev) * lp).sum()
        loss.backward()
        ref = gg[f"{mode}/gx"]
        err = np.abs(zt.grad.cpu().numpy() - ref).max(1) / np.abs(ref).max()
        ex = np.abs(x.detach().cpu().numpy() - gg["x64"]).max()
        worst = 0.0
        for n_, p in flow.named_parameters():
            for key in (f"{mode}/pg/{n_}", f"{mode}/pg_sample/{n_}"):
                if key in gg:
                    g = p.grad.cpu().numpy().reshape(-1).astype(np.float64)
                    if "sample" in key:
                        g = g[grad_sample_idx(g.size)]
                    worst = max(worst, float(np.abs(g - gg[key]).max() / max(np.abs(gg[key]).max(), 1e-30)))
        print(f"{name:18s} {mode:6s} x err {ex:.2e} | gz err: all rows {err.max():.2e}, smooth rows {err[~kink].max():.2e} ({int(kink.sum())} non-smooth) | params (all rows) {worst:.2e}")
