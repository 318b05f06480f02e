"""Per-tile timeline of the wide fused layer kernel (clock64 stamps of CTA 0, third tile).
Usage: python profiles/wide_timeline.py cfg3|cfg5"""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0)
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
if which == 'cfg2':
    flow, D, C = zuko.flows.NSF(16, 8, transforms=1, bins=8, hidden_features=[256] * 3), 16, 8
elif which == 'cfg3':
    flow, D, C = zuko.flows.MAF(32, 0, transforms=1, hidden_features=[512] * 4), 32, 0
else:
    flow, D, C = zuko.flows.NSF(64, 16, transforms=1, bins=16, hidden_features=[512] * 3), 64, 16
flow = flow.to(dev)
B = 1 << 20 if which == 'cfg2' else 1 << 19
x = torch.randn(B, D, device=dev); c = torch.randn(B, C, device=dev) if C else None
flow(c).log_prob(x); torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
E.lib().zk_debug_timeline(buf.data_ptr())
flow(c).log_prob(x); torch.cuda.synchronize()
E.lib().zk_debug_timeline(None)
t = buf.cpu().numpy()
t0 = t[0]
names = {0: 'epi: tile start', 1: 'epi: input staged', 2: 'epi: tile end'}
for l in range(4):
    for ch in range(4):
        names[8 + 8 * l + 2 * ch] = f'epi L{l}c{ch}: d_full seen'; names[9 + 8 * l + 2 * ch] = f'epi L{l}c{ch}: A written'
for ch in range(40):
    names[80 + 2 * ch] = f'epi out c{ch}: d_full seen'; names[81 + 2 * ch] = f'epi out c{ch}: dims done'
for i in range(256, 512): names[i] = f'mma: entry {i - 256} issued'
ev = sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t) if v != 0)
prev = None
for dt, n in ev:
    gap = '' if prev is None or not n.startswith('mma') else f'  (+{dt - prev})'
    if n.startswith('mma'): prev = dt
    print(f'{dt:8d}  {n}{gap}')
