"""Per-tile timeline of the dual-tile fused kernel on one cfg2 layer (clock64 stamps of CTA 0, third tile)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0); torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
flow = zuko.flows.NSF(16, 8, transforms=1, bins=8, hidden_features=[256] * 3).to(dev)
B = 1 << 20
x = torch.randn(B, 16, device=dev); c = torch.randn(B, 8, device=dev)
flow(c).log_prob(x); torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
E.lib().zk_debug_timeline(buf.data_ptr())
flow(c).log_prob(x); torch.cuda.synchronize()
E.lib().zk_debug_timeline(None)
t = buf.cpu().numpy()
t0 = min(v for v in (t[0], t[128]) if v)
names = {}
for u in range(2):
    b = 128 * u
    names[b] = f'G{u}: tile start'; names[b + 1] = f'G{u}: input staged'; names[b + 2] = f'G{u}: tile end'
    for l in range(4):
        for ch in range(2):
            names[b + 8 + 8 * l + 2 * ch] = f'G{u} L{l}c{ch}: d_full seen'; names[b + 9 + 8 * l + 2 * ch] = f'G{u} L{l}c{ch}: A written'
    for ch in range(20):
        names[b + 48 + 2 * ch] = f'G{u} out c{ch}: d_full seen'; names[b + 49 + 2 * ch] = f'G{u} out c{ch}: dims done'
for i in range(256, 320): names[i] = f'mma: entry {i - 256} issued'
for i in range(320, 384): names[i] = f'  weight scout: entry {i - 320} landed'
for i in range(384, 448): names[i] = f'  operand scout: entry {i - 384} ready'
for u in range(2):
    for l in range(4):
        for ch in range(2): names[128 * u + 100 + 2 * l + ch] = f'G{u} L{l}c{ch}: drained (d_empty arrive)'
for i in range(448, 512): names[i] = f'    producer: entry {i - 448} TMA issued (leader CTA)'
lat = [(i - 448, int(t[320 + i - 448] - t[i])) for i in range(448, 512) if t[i] and t[320 + i - 448]]
print('TMA issue -> landed (cycles):', ' '.join(f'{j}:{d}' for j, d in lat))
ev = sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t) if v != 0)
prev = None
for dt, n in ev:
    gap = '' if prev is None or not n.startswith('mma') else f'  (+{dt - prev})'
    if n.startswith('mma'): prev = dt
    print(f'{dt:8d}  {n}{gap}')
