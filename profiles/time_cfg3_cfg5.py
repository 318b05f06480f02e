"""Times BASELINE configs 3 and 5 (hidden width 512: per-layer tcgen05 GEMM path + stand-alone bijector kernel)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
dev = torch.device('cuda:0')
def run(name, flow, B, D, C):
    flow = flow.to(dev)
    x = torch.randn(B, D, device=dev); c = torch.randn(B, C, device=dev) if C else None
    flow(None if c is None else c[:4096]).log_prob(x[:4096]); torch.cuda.synchronize()
    d = flow(c); d.log_prob(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): lp = d.log_prob(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: B={B} {ms:.2f} ms/step -> {B / ms * 1e3:.3e} samples/s")
torch.manual_seed(0)
run("cfg3 MAF(32, T8, [512]*4) log_prob", zuko.flows.MAF(32, 0, transforms=8, hidden_features=[512] * 4), 1 << 20, 32, 0)
run("cfg5 NSF(64, 16, T8, K16, [512]*3) log_prob (one GPU's 2^21-row shard, run as 2^19)", zuko.flows.NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3), 1 << 19, 64, 16)
