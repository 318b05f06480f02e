"""Quick dead-lock / sanity probe of the fused layer kernel (run under `timeout`): a few shapes,
several tiles per CTA, compared with the per-layer path."""
import sys, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
dev = torch.device('cuda:0')
torch.manual_seed(0)
for name, ctor, C in [
    ('nsf16c8_h256', lambda: zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3), 8),
    ('nsf64_k16_h64', lambda: zuko.flows.NSF(64, 0, transforms=2, bins=16), 0),
    ('nsf5c3_h192', lambda: zuko.flows.NSF(5, 3, transforms=2, bins=8, hidden_features=[192, 192]), 3),
    ('maf70c100_h128', lambda: zuko.flows.MAF(70, 100, transforms=2, hidden_features=[128] * 2), 100),
]:
    flow = ctor().to(dev)
    D = flow.base.loc.shape[0]
    for B in (100, 70000):
        x = torch.randn(B, D, device=dev)
        c = torch.randn(B, C, device=dev) if C else None
        lp = flow(c).log_prob(x)
        torch.cuda.synchronize()
        prev = E.lib().zk_set_fused_layers(0)
        lp_u = flow(c).log_prob(x)
        E.lib().zk_set_fused_layers(prev)
        print(name, B, float((lp - lp_u).abs().max()), flush=True)
print('OK')
