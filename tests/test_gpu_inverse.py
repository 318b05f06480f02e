"""GPU tests of the dimension-sequential autoregressive inverse (ar_inverse.cu) against the
reference's sweep-based fixed point (run by the engine with zk_set_fused_layers(0)) and the
fp64 oracle (which restates the sweeps of zuko/transforms.py:994-1000 literally)."""

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from cases import rel_err
from oracle import oracle as O
from zuko_b200 import _engine as E

pytestmark = pytest.mark.gpu

FLOWS = {
    "nsf64_k16": lambda: zuko.flows.NSF(64, 0, transforms=2, bins=16),                              # BASELINE config 4 shape
    "nsf16c8_h256": lambda: zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3), # config 2 shape
    "maf5c2_randperm": lambda: zuko.flows.MAF(5, 2, transforms=3, randperm=True, hidden_features=[24]),
    "nsf5_passes2": lambda: zuko.flows.NSF(5, 0, transforms=2, passes=2, hidden_features=[32, 32]),
    "maf32_h512": lambda: zuko.flows.MAF(32, 0, transforms=2, hidden_features=[512] * 2),
    "nsf3c5": lambda: zuko.flows.NSF(3, 5, transforms=3),
}


@pytest.mark.parametrize("name", list(FLOWS))
@pytest.mark.parametrize("B", [1, 100, 3000])
def test_sequential_inverse_matches_sweeps_and_oracle(device, name, B):
    torch.manual_seed(21)
    flow_cpu = FLOWS[name]().eval()
    spec = O.flowspec_from_module(flow_cpu)
    D = flow_cpu.base.loc.shape[0]
    C = flow_cpu.transform.transforms[0].context
    g = torch.Generator().manual_seed(B + 1)
    z = torch.randn(B, D, generator=g)
    c = torch.randn(B, C, generator=g) if C else None
    flow = FLOWS[name]()
    flow.load_state_dict(flow_cpu.state_dict())
    flow = flow.to(device)
    zd, cd = z.to(device), (None if c is None else c.to(device))
    flow(cd).transform.inv(zd[:1])  # packs the layers (mask / split kernels) outside the count
    n0 = E.lib().zk_launch_count()
    x_fast = flow(cd).transform.inv(zd)
    launches = E.lib().zk_launch_count() - n0
    assert launches == len(flow.transform.transforms), launches  # ONE kernel per layer, not `passes` sweeps
    prev = E.lib().zk_set_fused_layers(0)
    try:
        x_sweeps = flow(cd).transform.inv(zd)
    finally:
        E.lib().zk_set_fused_layers(prev)
    ref = spec.inverse(z.numpy(), None if c is None else c.numpy())
    # fp32 FMA throughout: agreement with the fp64 fixed point at the fp32 conditioning of the map
    assert rel_err(x_fast.detach().cpu().numpy(), ref) < 2e-5
    assert rel_err(x_sweeps.detach().cpu().numpy(), ref) < 5e-5
    # round trip through the forward kernels (tests/test_flows.py:57-61: atol 1e-4)
    assert torch.allclose(flow(cd).transform(x_fast), zd, atol=1e-4)


def test_rsample_and_log_prob_uses_sequential_inverse(device):
    torch.manual_seed(5)
    flow = zuko.flows.NSF(64, 0, transforms=4, bins=16).to(device)
    dist = flow()
    x, lp = dist.rsample_and_log_prob((4096,))
    assert x.shape == (4096, 64) and torch.isfinite(x).all()
    assert torch.allclose(lp, dist.log_prob(x), rtol=1e-6, atol=1e-4)
