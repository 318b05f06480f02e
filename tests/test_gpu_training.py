"""GPU tests of the training-side plumbing around the engine's autograd seam: gradients on the
UNFUSED path (`base.log_prob(z) + ladj`), a trainable DiagNormal base, the in-place weight refresh an
optimizer step triggers (zk_layer_update_weights), and the guards that replace silent wrong answers."""

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from zuko_b200 import _engine as E
from zuko_b200.distributions import DiagNormal
from zuko_b200.flows import MaskedAutoregressiveTransform
from zuko_b200.lazy import Flow, UnconditionalDistribution

pytestmark = pytest.mark.gpu


def _grads(flow):
    return {n: p.grad.detach().clone() for n, p in flow.named_parameters() if p.grad is not None}


def test_unfused_path_carries_the_base_gradient(device):
    """Flow(single lazy transform, base) is not a ComposedTransform, so log_prob is
    `base.log_prob(z) + ladj` — the base term must back-propagate into z.  Same parameters wrapped in a
    one-member list take the fused engine call (checked against the reference's autograd goldens in
    test_gpu_backward): the two gradients must agree."""
    torch.manual_seed(3)
    t = MaskedAutoregressiveTransform(5, 3, hidden_features=[64, 64])
    base = lambda: UnconditionalDistribution(DiagNormal, torch.zeros(5), torch.ones(5), buffer=True)  # noqa: E731
    single = Flow(t, base()).to(device)
    fused = Flow([t], base()).to(device)  # shares the transform's parameters
    x = torch.randn(512, 5, device=device)
    c = torch.randn(512, 3, device=device)
    assert single(c)._flow_call() is None and fused(c)._flow_call() is not None
    (-single(c).log_prob(x).mean()).backward()
    g_single = [p.grad.detach().clone() for p in t.parameters()]  # the two flows share these parameters
    for p in t.parameters():
        p.grad = None
    (-fused(c).log_prob(x).mean()).backward()
    g_fused = [p.grad.detach().clone() for p in t.parameters()]
    assert len(g_single) == len(g_fused) > 0
    for k, (a, b) in enumerate(zip(g_single, g_fused)):
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() < 5e-5 * scale, k
    # and d/dx: the base term dominates it
    xg = x.clone().requires_grad_(True)
    single(c).log_prob(xg).sum().backward()
    xf = x.clone().requires_grad_(True)
    fused(c).log_prob(xf).sum().backward()
    assert torch.allclose(xg.grad, xf.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("fused", [True, False])
def test_trainable_base_gets_gradients(device, fused):
    """A parametric base (UnconditionalDistribution(DiagNormal, loc, scale), buffer=False: nn.Parameters,
    zuko/lazy.py:242-287) trains in the reference; here loc / scale join the autograd seam."""
    torch.manual_seed(4)
    t = MaskedAutoregressiveTransform(4, 0, hidden_features=[32])
    loc, scale = torch.randn(4) * 0.3, torch.rand(4) + 0.5
    flow = Flow([t] if fused else t, UnconditionalDistribution(DiagNormal, loc, scale)).to(device)
    names = dict(flow.named_parameters())
    base_params = [p for n, p in names.items() if n.startswith("base")]
    assert len(base_params) == 2
    x = torch.randn(300, 4, device=device)
    w = torch.rand(300, device=device)
    (flow().log_prob(x) * w).sum().backward()
    with torch.no_grad():
        z, _ = flow().transform.call_and_ladj(x)
    l2, s2 = [p.detach().clone().requires_grad_(True) for p in base_params]
    ref = (torch.distributions.Independent(torch.distributions.Normal(l2, s2), 1).log_prob(z) * w).sum()
    ref.backward()
    for p, r in zip(base_params, (l2, s2)):
        assert p.grad is not None
        assert torch.allclose(p.grad, r.grad, rtol=2e-4, atol=1e-4), (p.grad, r.grad)
    # rsample_and_log_prob: the explicit dependence of the log-density on loc / scale (+ through z = loc + eps scale)
    for p in flow.parameters():
        p.grad = None
    torch.manual_seed(0)
    xs, lp = flow().rsample_and_log_prob((64,))
    lp.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in base_params)


def test_optimizer_step_refreshes_the_pack_in_place(device):
    """An optimizer step moves only the weights' versions: the cached handle is refreshed by
    zk_layer_update_weights (same OwnedLayer, generation + 1) and every path — fused forward, per-layer
    path, dimension-sequential inverse, backward — sees the new weights: results equal a flow freshly
    built from the updated state dict, bit for bit."""
    torch.manual_seed(8)
    flow = zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3).to(device)
    x = torch.randn(2048, 16, device=device)
    c = torch.randn(2048, 8, device=device)
    opt = torch.optim.SGD(flow.parameters(), lr=1e-2)
    (-flow(c).log_prob(x).mean()).backward()
    refs = [t._zk_layer_ref() for t in flow.transform.transforms]
    with torch.no_grad():
        flow(c[:64]).transform.inv(x[:64])  # builds the inverse packs that the refresh must invalidate
    opt.step()
    opt.zero_grad()
    lp = flow(c).log_prob(x)  # refresh happens here
    assert all(t._zk_layer_ref() is r for t, r in zip(flow.transform.transforms, refs))
    assert [r.generation for r in refs] == [1, 1]
    fresh = zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3)
    fresh.load_state_dict(flow.state_dict())
    fresh = fresh.to(device)
    with torch.no_grad():
        assert torch.equal(lp.detach(), fresh(c).log_prob(x))
        assert torch.equal(flow(c[:300]).transform.inv(x[:300]), fresh(c[:300]).transform.inv(x[:300]))
        prev = E.lib().zk_set_fused_layers(0)
        try:
            assert torch.equal(flow(c).log_prob(x), fresh(c).log_prob(x))
        finally:
            E.lib().zk_set_fused_layers(prev)
    (-lp.mean()).backward()
    (-fresh(c).log_prob(x).mean()).backward()
    for (n, p), (_, q) in zip(flow.named_parameters(), fresh.named_parameters()):
        assert torch.allclose(p.grad, q.grad, rtol=1e-6, atol=1e-9), n


def test_backward_after_an_in_place_refresh_raises(device):
    torch.manual_seed(9)
    flow = zuko.flows.MAF(6, 0, transforms=2, hidden_features=[64]).to(device)
    x = torch.randn(256, 6, device=device)
    loss = -flow().log_prob(x).mean()
    with torch.no_grad():
        for p in flow.parameters():
            p.add_(0.01)  # bumps _version like an optimizer step
    flow().log_prob(x)  # refreshes the packs in place
    with pytest.raises(RuntimeError, match="refreshed in place"):
        loss.backward()


def test_invalidate_after_data_writes(device):
    """Writes through .data do not bump _version: the stale pack is served until
    zuko_b200.invalidate() is called."""
    torch.manual_seed(10)
    flow = zuko.flows.MAF(6, 0, transforms=1, hidden_features=[64]).to(device)
    x = torch.randn(128, 6, device=device)
    with torch.no_grad():
        a = flow().log_prob(x)
        for p in flow.parameters():
            p.data.mul_(0.5)
        stale = flow().log_prob(x)
        assert torch.equal(a, stale)  # documented behaviour of the (data_ptr, _version) key
        zuko.invalidate(flow)
        b = flow().log_prob(x)
    assert not torch.equal(a, b)


def test_standalone_conditioner_refuses_to_drop_gradients(device):
    net = zuko.nn.MLP(4, 3, [32]).to(device)
    x = torch.randn(8, 4, device=device)
    with pytest.raises(NotImplementedError, match="forward-only"):
        net(x.clone().requires_grad_(True))
    zuko.nn._EngineMLP._warned_detached = False
    with pytest.warns(RuntimeWarning, match="NOT connected"):
        y = net(x)  # parameters still carry requires_grad: allowed for inference, said out loud once
    assert not y.requires_grad
    with torch.no_grad():
        assert net(x).shape == (8, 3)


def test_engine_calls_on_two_streams_do_not_share_scratch(device):
    torch.manual_seed(12)
    flow = zuko.flows.NSF(16, 8, transforms=2, bins=8, hidden_features=[256] * 3).to(device)
    x = torch.randn(1 << 15, 16, device=device)
    c = torch.randn(1 << 15, 8, device=device)
    with torch.no_grad():
        ref = flow(c).log_prob(x)
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
        outs = []
        for _ in range(4):
            with torch.cuda.stream(s1):
                a = flow(c).log_prob(x)
            with torch.cuda.stream(s2):
                b = flow(c).transform.inv(x)  # a different footprint in the same scratch, were it shared
            outs.append(a)
        torch.cuda.synchronize()
    assert all(torch.equal(o, ref) for o in outs)
    assert np.isfinite(b.cpu().numpy()).all()


def test_fast_and_ieee_bijector_gradients_agree(device):
    """zk_set_fast_math switches the RQS backward pair math between MUFU reciprocal / ex2 and IEEE division / expf
    (bijector_grad.cuh); both are held to the reference's gradients elsewhere — here: they agree with each other
    far inside that bar."""
    torch.manual_seed(14)
    flow = zuko.flows.NSF(6, 2, transforms=2, bins=8, hidden_features=[64, 64]).to(device)
    x = torch.randn(4096, 6, device=device)
    c = torch.randn(4096, 2, device=device)
    grads = {}
    for fast in (1, 0):
        prev = E.lib().zk_set_fast_math(fast)
        try:
            for p in flow.parameters():
                p.grad = None
            (-flow(c).log_prob(x).mean()).backward()
            grads[fast] = [p.grad.detach().clone() for p in flow.parameters()]
        finally:
            E.lib().zk_set_fast_math(prev)
    for a, b in zip(grads[1], grads[0]):
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() < 1e-5 * scale
