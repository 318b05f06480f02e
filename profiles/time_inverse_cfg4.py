import sys, time, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0)
dev = torch.device('cuda:0')
flow = zuko.flows.NSF(64, 0, transforms=8, bins=16).to(dev)  # BASELINE config 4
for N, fused in ((1 << 20, 1), (1 << 16, 0)):
    E.lib().zk_set_fused_layers(fused)
    z = torch.randn(N, 64, device=dev)
    t = flow().transform
    t.inv(z[:1024]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x = t.inv(z); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"cfg4 inverse N={N} {'sequential kernel' if fused else 'reference sweeps (64 x 8)'}: {ms:.1f} ms -> {N / ms * 1e3:.3e} samples/s")
E.lib().zk_set_fused_layers(1)
