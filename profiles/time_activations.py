"""MAF(32, T8, [512]^4) log_prob at 2^20 rows with ReLU / ELU / SiLU / GELU conditioners (fused pair kernel)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
B = 1 << 20
x = torch.randn(B, 32, device=dev)
base = None
for name, act in (("ReLU", None), ("ELU", torch.nn.ELU), ("SiLU", torch.nn.SiLU), ("GELU", torch.nn.GELU), ("Tanh", torch.nn.Tanh)):
    torch.manual_seed(0)
    kw = {} if act is None else {"activation": act}
    flow = zuko.flows.MAF(32, 0, transforms=8, hidden_features=[512] * 4, **kw).to(dev)
    d = flow(); d.log_prob(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.log_prob(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    base = base or ms
    print(f"cfg3 shape, {name:5s}: {ms:.2f} ms/step ({ms / base:.2f} x ReLU)")
