"""Times one training step (log_prob forward + backward through zk_flow_backward) of BASELINE
config 2 with CUDA events.  Usage: python profiles/time_backward.py [rows]"""

import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import zuko_b200 as zuko  # noqa: E402
from zuko_b200 import _engine as E  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
ONCE = len(sys.argv) > 2 and sys.argv[2] == "once"  # one warm-up + one step (for an ncu launch list)
dev = torch.device("cuda:0")
torch.manual_seed(0)
flow = zuko.flows.NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3).to(dev)
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 16, generator=g).to(dev)
c = torch.randn(B, 8, generator=g).to(dev)


def step():
    for p in flow.parameters():
        p.grad = None
    loss = -flow(c).log_prob(x).mean()
    loss.backward()
    return loss


for _ in range(1 if ONCE else 3):
    step()
torch.cuda.synchronize()
if ONCE:
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
n0 = E.lib().zk_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
with torch.no_grad():
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(iters):
        flow(c).log_prob(x)
    f1.record()
    torch.cuda.synchronize()
fwd = f0.elapsed_time(f1) / iters
dims = [24, 256, 256, 256, 368]
flops = 2.0 * sum(a * b for a, b in zip(dims[:-1], dims[1:])) * B * 4
print(json.dumps({"workload": "NSF(16,8,T4,K8,[256]^3) training step (forward + backward)",
                  "rows": B, "ms_per_step": ms, "ms_forward_only": fwd, "samples_per_s": B / (ms * 1e-3),
                  "launches_per_step": (E.lib().zk_launch_count() - n0) / (2 * iters) * 2,
                  "backward_tflops_algorithmic": 3 * flops / ((ms - fwd) * 1e-3) / 1e12}))
