"""cfg2 / cfg3 on the PER-LAYER path (fused kernels off): linear_tc_kernel + uni_kernel per flow layer; CUDA-event ms."""
import sys
import torch
sys.path.insert(0, "/root/repo")
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0); torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
E.lib().zk_set_fused_layers(0)
for name, make, B, D, C in (("cfg2", lambda: zuko.flows.NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 1 << 20, 16, 8),
                            ("cfg3", lambda: zuko.flows.MAF(32, 0, transforms=8, hidden_features=[512] * 4), 1 << 20, 32, 0)):
    flow = make().to(dev)
    x = torch.randn(B, D, device=dev); c = torch.randn(B, C, device=dev) if C else None
    for _ in range(3): flow(c).log_prob(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): flow(c).log_prob(x)
    e1.record(); torch.cuda.synchronize()
    print(f"{name} per-layer path: {e0.elapsed_time(e1) / 5:.3f} ms per log_prob of {B} rows", flush=True)
