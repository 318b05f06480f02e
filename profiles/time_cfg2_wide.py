"""cfg2 (NSF(16, 8, T4, K8, [256]^3) log_prob, 2^20 rows) on the three fused kernels: dual-tile CTA pairs
(default), one-tile CTA pairs (zk_set_dual_tiles(0)), one CTA per tile (+ zk_set_wide_min_hidden(384))."""
import sys, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
B = 1 << 20
torch.manual_seed(1)
x = torch.randn(B, 16, device=dev); c = torch.randn(B, 8, device=dev)
outs = {}
for name, dual, min_h in (("narrow (one CTA per tile)", 0, 384), ("pair, one tile", 0, 256), ("pair, two sub-tiles", 1, 256)):
    E.lib().zk_set_dual_tiles(dual); E.lib().zk_set_wide_min_hidden(min_h)
    torch.manual_seed(0)
    flow = zuko.flows.NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3).to(dev)
    d = flow(c); lp = d.log_prob(x); torch.cuda.synchronize()
    for _ in range(5): d.log_prob(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lp = d.log_prob(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    outs[name] = lp
    print(f"cfg2 log_prob B=2^20, kernel = {name}: {ms:.3f} ms/step -> {B / ms * 1e3:.4e} samples/s", flush=True)
ref = outs["narrow (one CTA per tile)"]
for k, v in outs.items(): print(f"max |lp[{k}] - lp[narrow]| = {(v - ref).abs().max().item():.3e}")
