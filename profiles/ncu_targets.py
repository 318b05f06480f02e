"""Launches ONE kernel class twice (warm-up + the launch ncu captures with `-s 1 -c 1`).
Usage: python profiles/ncu_targets.py cfg2|cfg3|cfg5|cfg4|rqs16
  cfg2  fused_layer_kernel<RQS,8>   one layer of NSF(16, 8, K8, [256]^3),   2^20 rows
  cfg3  fused_wide_kernel<AFFINE>   one layer of MAF(32, [512]^4),          2^20 rows
  cfg5  fused_wide_kernel<RQS,16>   one layer of NSF(64, 16, K16, [512]^3), 2^19 rows
  cfg4  ar_inverse_kernel<RQS,16>   one layer of NSF(64, K16, [64, 64]),    2^18 rows
  rqs16 uni_kernel<RQS,16>          stand-alone bijector, D = 64, K = 16,   2^19 rows (phi in HBM)"""
import sys, torch
sys.path.insert(0, '/root/repo')
import zuko_b200 as zuko
from zuko_b200 import _engine as E
torch.manual_seed(0); torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
which = sys.argv[1]
if which == 'rqs16':
    B, D, K = 1 << 19, 64, 16
    P = 3 * K - 1
    x = torch.randn(B, D, device=dev); phi = torch.randn(B, D * P, device=dev); y = torch.empty_like(x); ladj = torch.zeros(B, device=dev)
    for _ in range(2):
        E.check(E.lib().zk_rqs_forward(x.data_ptr(), D, phi.data_ptr(), D * P, B, D, K, 5.0, 1e-3, y.data_ptr(), D, ladj.data_ptr(), 0, E.stream_ptr(dev)))
    torch.cuda.synchronize()
    sys.exit(0)
cfg = {'cfg2': (lambda: zuko.flows.NSF(16, 8, transforms=1, bins=8, hidden_features=[256] * 3), 16, 8, 1 << 20),
       'cfg3': (lambda: zuko.flows.MAF(32, 0, transforms=1, hidden_features=[512] * 4), 32, 0, 1 << 20),
       'cfg5': (lambda: zuko.flows.NSF(64, 16, transforms=1, bins=16, hidden_features=[512] * 3), 64, 16, 1 << 19),
       'cfg4': (lambda: zuko.flows.NSF(64, 0, transforms=1, bins=16), 64, 0, 1 << 18)}[which]
flow = cfg[0]().to(dev)
D, C, B = cfg[1:]
x = torch.randn(B, D, device=dev); c = torch.randn(B, C, device=dev) if C else None
for _ in range(2):
    if which == 'cfg4': flow(c).transform.inv(x)
    else: flow(c).log_prob(x)
torch.cuda.synchronize()
