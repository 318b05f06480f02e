"""Diagnostic: per-row error of d/dx (tensor-core vs fp32 backward vs fp64 reference autograd)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import build_flow, grad_inputs  # noqa: E402
from zuko_b200 import _engine as E  # noqa: E402

dev = torch.device("cuda:0")
for name in sys.argv[1:] or ["cfg2_nsf"]:
    gg, x, c = grad_inputs(name)
    flow = build_flow(name).to(dev)
    res = {}
    for mode in (1, 0):
        prev = E.lib().zk_set_tc_backward(mode)
        for p in flow.parameters():
            p.grad = None
        xt = torch.from_numpy(x).to(dev).requires_grad_()
        ct = None if c is None else torch.from_numpy(c).to(dev).requires_grad_()
        lp = flow(ct).log_prob(xt)
        (torch.from_numpy(gg["g"]).float().to(dev) * lp).sum().backward()
        E.lib().zk_set_tc_backward(prev)
        res[mode] = (xt.grad.cpu().numpy().astype(np.float64), {n: p.grad.cpu().numpy().astype(np.float64).reshape(-1) for n, p in flow.named_parameters()})
    ref = gg["lp/gx"]
    scale = np.abs(ref).max()
    for mode, tag in ((1, "tc  "), (0, "fp32")):
        err = np.abs(res[mode][0] - ref).max(axis=1) / scale
        order = np.argsort(-err)[:5]
        print(name, tag, "gx max err %.3e | rows > 1e-5: %d of %d | median %.2e | worst rows" % (err.max(), (err > 1e-5).sum(), err.size, np.median(err)), [(int(i), float("%.2e" % err[i])) for i in order])
    # parameter gradients: tc vs fp32 relative to the largest entry
    worst = []
    for n in res[0][1]:
        a, b = res[1][1][n], res[0][1][n]
        worst.append((float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)), n))
    worst.sort(reverse=True)
    print(name, "param grads tc vs fp32, worst:", [("%.2e" % e, n) for e, n in worst[:4]])
