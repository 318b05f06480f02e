"""Diagnostic: cfg4 (NSF 64, K16, hidden [64,64]) at B=1500 — forward chain and gradient vs the oracle."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cases import build_flow  # noqa: E402
from oracle import oracle as O, oracle_grad as OG  # noqa: E402
from zuko_b200 import _engine as E  # noqa: E402

dev = torch.device("cuda:0")
name, B = "cfg4_nsf", int(sys.argv[1]) if len(sys.argv) > 1 else 1500
flow = build_flow(name).to(dev)
spec = O.flowspec_from_module(build_flow(name))
gen = torch.Generator().manual_seed(11)
x = torch.randn(B, 64, generator=gen).numpy()
g = torch.randn(B, generator=gen).numpy()
n = 64
xt = torch.from_numpy(x).to(dev)
with torch.no_grad():
    lp = flow().log_prob(xt).cpu().numpy().astype(np.float64)
    z, ladj = flow().transform.call_and_ladj(xt)
ref_lp = spec.log_prob(x[:n])
rz, rl = spec.forward(x[:n])
print("log_prob err", np.abs(lp[:n] - ref_lp).max(), "| z err", np.abs(z.cpu().numpy()[:n] - rz).max(), "| ladj err", np.abs(ladj.cpu().numpy()[:n] - rl).max())
for fused in (1, 0):
    prev = E.lib().zk_set_fused_layers(fused)
    xg = torch.from_numpy(x).to(dev).requires_grad_()
    (torch.from_numpy(g).to(dev) * flow().log_prob(xg)).sum().backward()
    E.lib().zk_set_fused_layers(prev)
    ogx, _, _ = OG.flow_backward(spec, x[:n], None, g_log_prob=g[:n])
    err = np.abs(xg.grad.cpu().numpy()[:n].astype(np.float64) - ogx)
    rows = err.max(1)
    print("fused" if fused else "unfused", "gx err max %.3e; rows > 1e-3:" % err.max(), np.nonzero(rows > 1e-3)[0][:20], "dims of worst row:", np.nonzero(err[rows.argmax()] > 1e-3)[0][:20])
