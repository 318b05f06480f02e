"""A/B of the pair kernels' MMA issue order (zk_set_mma_order 0 / 1 / 2) on cfg2, cfg3, cfg5: CUDA-event time per
log_prob call, interleaved over the three orders so that clock drift hits them equally; also checks the three
orders agree to fp32 rounding.  Usage: python profiles/time_mma_order.py"""
import sys

import torch

sys.path.insert(0, "/root/repo")
import zuko_b200 as zuko  # noqa: E402
from zuko_b200 import _engine as E  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
CFG = {
    "cfg2": (lambda: zuko.flows.NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 1 << 20, 16, 8),
    "cfg3": (lambda: zuko.flows.MAF(32, 0, transforms=8, hidden_features=[512] * 4), 1 << 20, 32, 0),
    "cfg5": (lambda: zuko.flows.NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3), 1 << 19, 64, 16),
}
L = E.lib()
for name, (make, B, D, C) in CFG.items():
    flow = make().to(dev)
    x = torch.randn(B, D, device=dev)
    c = torch.randn(B, C, device=dev) if C else None
    outs = {}
    for order in (0, 1, 2):
        L.zk_set_mma_order(order)
        outs[order] = flow(c).log_prob(x).clone()
    torch.cuda.synchronize()
    d01 = (outs[0] - outs[1]).abs().max().item()
    d02 = (outs[0] - outs[2]).abs().max().item()
    times = {0: [], 1: [], 2: []}
    for rep in range(6):
        for order in (0, 1, 2):
            L.zk_set_mma_order(order)
            flow(c).log_prob(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                flow(c).log_prob(x)
            e1.record()
            torch.cuda.synchronize()
            times[order].append(e0.elapsed_time(e1) / 5)
    L.zk_set_mma_order(2)
    line = "  ".join(f"order {o}: median {sorted(t)[len(t) // 2]:.3f} ms (min {min(t):.3f}, max {max(t):.3f})" for o, t in times.items())
    print(f"{name} B={B}: {line}   max |lp0 - lp1| {d01:.2e}  |lp0 - lp2| {d02:.2e}  (|lp| ~ {outs[0].abs().mean().item():.1f})", flush=True)
