import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0)
dev = torch.device('cuda:0')
flow = zuko.flows.NSF(16, 8, transforms=1, bins=8, hidden_features=[256]*3)
if len(sys.argv) > 1:
    flow.transform.transforms[0].hyper.gemm_mode = sys.argv[1]
flow = flow.to(dev)
B = 1 << 20
x = torch.randn(B, 16, device=dev); c = torch.randn(B, 8, device=dev)
flow(c).log_prob(x); torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
E.lib().zk_debug_timeline(buf.data_ptr())
flow(c).log_prob(x); torch.cuda.synchronize()
E.lib().zk_debug_timeline(None)
t = buf.cpu().numpy()
t0 = t[48]
names = {48: 'epi: tile start', 49: 'epi: input staged', 50: 'epi: tile end', 51: 'epi: in_full seen', 52: 'epi: staging stores issued', 56: 'in-producer: waiting in_empty (tile 2)', 57: 'in-producer: issuing load (tile 2)'}
for l in range(4):
    names[8*l+0] = f'mma L{l}: a_ready[0] seen'; names[8*l+1] = f'mma L{l}: a_ready[last] seen'; names[8*l+2] = f'mma L{l}: first W stage full'; names[8*l+3] = f'mma L{l}: all issued'
for l in range(3):
    for ch in range(2):
        b = 64 + 16*l + 4*ch
        names[b] = f'epi L{l}c{ch}: d_full seen'; names[b+1] = f'epi L{l}c{ch}: computed'; names[b+2] = f'epi L{l}c{ch}: a_free seen'; names[b+3] = f'epi L{l}c{ch}: A written'
for ch in range(8):
    names[160+2*ch] = f'epi last c{ch}: d_full seen'; names[161+2*ch] = f'epi last c{ch}: dims done'
for kb in range(4):
    names[200+3*kb] = f'mma L1c1 kb{kb}: waiting w_full'; names[201+3*kb] = f'mma L1c1 kb{kb}: w_full seen'; names[202+3*kb] = f'mma L1c1 kb{kb}: issued+committed'
    names[220+3*kb] = f'producer L1c1 kb{kb}: waiting w_empty'; names[221+3*kb] = f'producer L1c1 kb{kb}: w_empty seen'; names[222+3*kb] = f'producer L1c1 kb{kb}: TMA issued'
for l in range(4):
    names[100+l] = f'mma L{l}: reached layer start'; names[104+l] = f'mma L{l}: d_empty seen'
for l in range(3):
    for ch in range(2): names[110+2*l+ch] = f'epi(warp 19) L{l}c{ch}: A written'
names[240]='epi L1c1: LDTM + wait::ld done'; names[241]='epi L1c1: d_empty arrived'; names[242]='epi L1c1: STTM issued'; names[243]='epi L1c1: wait::st done'
for i in range(256, 512): names[i] = f'mma: schedule entry {i - 256} issued + committed'
ev = sorted((int(v - t0), names.get(i, str(i))) for i, v in enumerate(t) if v != 0)
for dt, n in ev: print(f'{dt:8d}  {n}')
