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
    # |x - x_ref| <= 1e-5 max(1, |x|) for the production (dimension-sequential) inverse; the
    # sweep path re-rounds the conditioner `passes` times and is held to the looser fp32 conditioning bound
    assert rel_err(x_fast.detach().cpu().numpy(), ref) < 1e-5
    assert rel_err(x_sweeps.detach().cpu().numpy(), ref) < 5e-5
    # round trip through the forward kernels (tests/test_flows.py:57-61: atol 1e-4)
    assert torch.allclose(flow(cd).transform(x_fast), zd, atol=1e-4)


RSLP_FLOWS = {
    **FLOWS,
    "ncsf4c3": lambda: zuko.flows.NCSF(4, 3, transforms=2),                                       # circular spline, BoxUniform base
    "nice5c3": lambda: zuko.flows.NICE(5, 3, transforms=3),                                       # coupling layers: per-layer ladj fallback
    "nsf6_elu": lambda: zuko.flows.NSF(6, 0, transforms=2, hidden_features=[64, 64], activation=torch.nn.ELU),      # non-ReLU conditioner on the sequential kernel
    "maf5_res_sweeps": lambda: zuko.flows.MAF(5, 0, transforms=2, hidden_features=[32, 32], residual=True),          # residual blocks: sweep inverse + per-layer ladj
}


@pytest.mark.parametrize("name", list(RSLP_FLOWS))
def test_inverse_and_log_prob_single_sweep_vs_oracle(device, name):
    """`rsample_and_log_prob` (distributions.py:129-138) for a supplied z: x against the oracle's
    inverse, the log-density against the oracle's fp64 log_prob AT THE ORACLE'S x (the quantity the
    reference returns), and — for flows made of dimension-sequential layers — launches == T: the
    ladj and the base log-density come out of the inverse sweep itself, no forward pass follows."""
    torch.manual_seed(33)
    flow_cpu = RSLP_FLOWS[name]().eval()
    spec = O.flowspec_from_module(flow_cpu)
    D = flow_cpu.transform.transforms[0].features if hasattr(flow_cpu.transform.transforms[0], "features") else flow_cpu.base.loc.shape[0]
    C = getattr(flow_cpu.transform.transforms[0], "context", 0)
    B = 777
    g = torch.Generator().manual_seed(9)
    if name.startswith("ncsf"):
        z = (torch.rand(B, D, generator=g) * 2 - 1) * 3.1
    else:
        z = torch.randn(B, D, generator=g)
    c = torch.randn(B, C, generator=g) if C else None
    flow = RSLP_FLOWS[name]()
    flow.load_state_dict(flow_cpu.state_dict())
    flow = flow.to(device)
    zd, cd = z.to(device), (None if c is None else c.to(device))
    with torch.no_grad():
        call, ctx = flow(cd)._flow_call()
        call.inverse(zd[:1], None if ctx is None else ctx[:1], with_log_prob=True)  # packs
        n0 = E.lib().zk_launch_count()
        x, lp = call.inverse(zd, ctx, with_log_prob=True)
        launches = E.lib().zk_launch_count() - n0
    cn = None if c is None else c.numpy()
    x_ref = spec.inverse(z.numpy(), cn)
    lp_ref = spec.log_prob(x_ref, cn)
    assert rel_err(x.cpu().numpy(), x_ref) < (5e-5 if "sweeps" in name else 1e-5)
    assert rel_err(lp.cpu().numpy(), lp_ref) < 1e-5
    if name in FLOWS or name.startswith("ncsf") or name == "nsf6_elu":
        assert launches == len(flow.transform.transforms) + (1 if name.startswith("ncsf") else 0), launches


def test_rsample_and_log_prob_uses_sequential_inverse(device):
    torch.manual_seed(5)
    flow = zuko.flows.NSF(64, 0, transforms=4, bins=16).to(device)
    dist = flow()
    x, lp = dist.rsample_and_log_prob((4096,))
    assert x.shape == (4096, 64) and torch.isfinite(x).all()
    # self-consistency only (the strict check against the oracle is test_inverse_and_log_prob_single_sweep_vs_oracle):
    # log_prob(x) is evaluated at the ENGINE's x, which carries ~1e-5 |x| of inverse error times |d log p / dx|
    assert rel_err(lp.detach().cpu().numpy(), dist.log_prob(x).detach().cpu().numpy()) < 1e-4


def test_lazy_inverse_members_run_inside_the_engine_call(device):
    """IAF-style flow (zuko/lazy.py:81-98): every member is the INVERSE of a masked autoregressive layer,
    so the density needs the sequential solve and sampling is one cheap forward.  Without autograd the
    whole `log_prob` is T launches (the sequential kernel emits minus the layer's ladj and, on the last
    member, the base density of its output); with autograd the call goes to the unfused chain.  Checked
    against the oracle: log p(x) = N(u; 0, I) - sum_l ladj_l(u_l) with u = the layers' inverses of x."""
    from zuko_b200.flows import MaskedAutoregressiveTransform
    from zuko_b200.lazy import Flow, LazyInverse, UnconditionalDistribution
    from zuko_b200.distributions import DiagNormal

    torch.manual_seed(41)
    D, C, T = 6, 3, 3
    layers = [MaskedAutoregressiveTransform(D, C, hidden_features=[64, 64], order=(torch.arange(D) if i % 2 == 0 else torch.arange(D).flip(0)))
              for i in range(T)]
    maf_cpu = Flow([t for t in layers], UnconditionalDistribution(DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)).eval()
    spec = O.flowspec_from_module(maf_cpu)  # the same layers, NOT inverted: the oracle composes them by hand below
    g = torch.Generator().manual_seed(2)
    x, c = torch.randn(500, D, generator=g), torch.randn(500, C, generator=g)
    # oracle: walk the inverted members forward: u_{l+1} = layer_l^{-1}(u_l), ladj_member = - ladj_layer(u_{l+1})
    u = x.numpy().astype(np.float64)
    total = np.zeros(500)
    for layer in spec.layers:
        u = layer.inverse(u, c.numpy(), np.float64)
        _, ladj = layer.forward(u, c.numpy(), np.float64)
        total -= ladj
    ref = total - 0.5 * (u**2).sum(-1) - 0.5 * D * np.log(2 * np.pi)
    iaf = Flow([LazyInverse(t) for t in layers], UnconditionalDistribution(DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)).to(device)
    xd, cd = x.to(device), c.to(device)
    with torch.no_grad():
        dist = iaf(cd)
        assert dist._flow_call() is not None and dist._flow_call()[0]._inverted == [True] * T
        lp = dist.log_prob(xd)
        n0 = E.lib().zk_launch_count()
        lp = dist.log_prob(xd)
        assert E.lib().zk_launch_count() - n0 == T
        z, ladj = dist.transform.call_and_ladj(xd)
        back = dist.transform.inv(z)  # sampling direction: the layers' forward kernels
    assert rel_err(lp.cpu().numpy(), ref) < 1e-5
    assert rel_err(z.cpu().numpy(), u) < 1e-5
    assert torch.allclose(back, xd, atol=1e-4)
    # with autograd the engine call (forward-only for inverted members) is bypassed: same value, gradients flow
    lp_g = iaf(cd).log_prob(xd)
    assert lp_g.requires_grad and rel_err(lp_g.detach().cpu().numpy(), ref) < 1e-5
    lp_g.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in iaf.parameters())
