"""GPU parity tests: the CUDA engine (through the public API -> C ABI) against the golden
vectors of the reference and against the oracle, plus size-independent properties.

Bar (BASELINE.json): log_prob within 1e-5 relative of the reference (fp32), permutation
indices bit-exact.  Tolerances are written next to each check.
"""

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from cases import BIG_CASES, FLOW_CASES, SMALL_CASES, assert_log_prob_parity, build_flow, load, rel_err
from oracle import oracle as O
from zuko_b200 import _engine as E
from zuko_b200.transforms import (
    MonotonicAffineTransform,
    MonotonicRQSTransform,
    PermutationTransform,
    RotationTransform,
    SoftclipTransform,
)

pytestmark = pytest.mark.gpu
U = load("units")


def dev_t(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def cpu(t):
    return t.detach().cpu().numpy().astype(np.float64)


@pytest.fixture(params=[1, 0], ids=["fast", "ieee"])
def math_mode(request):
    prev = E.lib().zk_set_fast_math(request.param)
    yield request.param
    E.lib().zk_set_fast_math(prev)


# --------------------------------------------------------------------------- #
# stand-alone bijectors (tests/test_transforms.py:12-121 of the reference)
# --------------------------------------------------------------------------- #


@pytest.mark.parametrize("tag,tol_y,tol_l", [("s01", 5e-6, 2e-5), ("s1", 1e-4, 5e-4), ("s3", 3e-3, 2e-2)])
def test_rqs_forward_inverse(device, math_mode, tag, tol_y, tol_l):
    """Per-element y / ladj vs the fp64 reference.  The spline is ill-conditioned in fp32
    once it is sharp: the tolerance is tied to the reference's own
    fp32-vs-fp64 deviation on the same inputs (x3 margin) with the stated floor."""
    K = 8
    phi, x = U[f"rqs_{tag}_phi"], U[f"rqs_{tag}_x"]
    p = dev_t(phi, device)
    t = MonotonicRQSTransform(p[..., :K], p[..., K : 2 * K], p[..., 2 * K :])
    y, ladj = t.call_and_ladj(dev_t(x, device))
    ref_dev_y = np.abs(U[f"rqs_{tag}_y32"].astype(np.float64) - U[f"rqs_{tag}_y64"])
    ref_dev_l = np.abs(U[f"rqs_{tag}_ladj32"].astype(np.float64) - U[f"rqs_{tag}_ladj64"])
    finite = np.isfinite(U[f"rqs_{tag}_ladj64"])
    ey = np.abs(cpu(y) - U[f"rqs_{tag}_y64"])[finite]
    el = np.abs(cpu(ladj) - U[f"rqs_{tag}_ladj64"])[finite]
    scale_y = np.maximum(np.abs(U[f"rqs_{tag}_y64"]), 1.0)[finite]
    assert np.all(ey <= np.maximum(tol_y * scale_y, 3 * ref_dev_y[finite])), ey.max()
    assert np.all(el <= np.maximum(tol_l, 3 * ref_dev_l[finite])), el.max()
    # inverse on fresh codomain points
    xi = t.inv(dev_t(U[f"rqs_{tag}_yq"], device))
    ref = U[f"rqs_{tag}_inv_of_yq64"]
    ref_dev = np.abs(U[f"rqs_{tag}_inv_of_yq32"].astype(np.float64) - ref)
    ex = np.abs(cpu(xi) - ref)
    assert np.all(ex <= np.maximum(tol_y * 4 * np.maximum(np.abs(ref), 1.0), 3 * ref_dev)), ex.max()


def test_rqs_edge_semantics(device, math_mode):
    """x = -5 (strict '<' puts it outside), x = 5, |x| > 5 (even 1e30): identity, ladj 0."""
    K = 8
    p = dev_t(U["rqs_s01_phi"], device)
    t = MonotonicRQSTransform(p[..., :K], p[..., K : 2 * K], p[..., 2 * K :])
    x = dev_t(U["rqs_s01_x"], device)
    y, ladj = t.call_and_ladj(x)
    for row in (0, 2, 4, 5, 9):
        assert y[row, 0].item() == x[row, 0].item(), row
        assert ladj[row, 0].item() == 0.0, row
    # x = +5 sits on the last knot (5 or 4.9999995 depending on the rounding of the cumsum):
    # identity up to one ulp either way, ladj = log(d_K) = 0
    assert abs(y[1, 0].item() - 5.0) <= 1e-6 and abs(ladj[1, 0].item()) <= 1e-5
    bad = torch.tensor([float("inf"), float("-inf"), float("nan")], device=device)
    t1 = MonotonicRQSTransform(p[0, 0, :K], p[0, 0, K : 2 * K], p[0, 0, 2 * K :])
    yb, lb = t1.call_and_ladj(bad)
    assert torch.isinf(yb[0]) and torch.isinf(yb[1]) and torch.isnan(yb[2])
    assert torch.isnan(lb).all()  # the reference yields NaN ladj for non-finite inputs


@pytest.mark.parametrize("K", [16, 5])
def test_rqs_shared_table_and_generic_bins(device, K):
    """Unbatched parameters (one shared table) incl. K = 5, which takes the runtime-K path."""
    p = dev_t(U[f"rqs_shared{K}_phi"], device)
    t = MonotonicRQSTransform(p[..., :K], p[..., K : 2 * K], p[..., 2 * K :])
    x = dev_t(U[f"rqs_shared{K}_x"], device)
    y, ladj = t.call_and_ladj(x)
    assert rel_err(cpu(y), U[f"rqs_shared{K}_y64"]) < 2e-5
    assert np.max(np.abs(cpu(ladj) - U[f"rqs_shared{K}_ladj64"])) < 1e-4
    assert rel_err(cpu(t.inv(x)), U[f"rqs_shared{K}_xinv64"]) < 1e-4
    assert torch.allclose(t.inv(y), x, atol=1e-4)  # tests/test_transforms.py:48


def test_affine(device, math_mode):
    p = dev_t(U["affine_phi"], device)
    t = MonotonicAffineTransform(p[..., 0], p[..., 1])
    x = dev_t(U["affine_x"], device)
    y, ladj = t.call_and_ladj(x)
    assert rel_err(cpu(y), U["affine_y64"]) < 5e-6
    assert np.max(np.abs(cpu(ladj) - U["affine_ladj64"])) < 5e-6
    assert rel_err(cpu(t.inv(x)), U["affine_xinv64"]) < 5e-6


@pytest.mark.parametrize("bound", [1, 11])
def test_softclip(device, bound):
    t = SoftclipTransform(bound=float(bound))
    x = dev_t(U["softclip_x"], device)
    y, ladj = t.call_and_ladj(x)
    assert rel_err(cpu(y), U[f"softclip{bound}_y64"]) < 2e-6
    assert np.max(np.abs(cpu(ladj) - U[f"softclip{bound}_ladj64"])) < 2e-6
    assert rel_err(cpu(t.inv(y)), U[f"softclip{bound}_xinv64"]) < 2e-5


def test_permutation_bit_exact(device):
    t = PermutationTransform(dev_t(U["perm_order"], device))
    x = dev_t(U["perm_x"], device)
    y, ladj = t.call_and_ladj(x)
    assert np.array_equal(y.cpu().numpy(), U["perm_y"])
    assert np.array_equal(t.inv(x).cpu().numpy(), U["perm_xinv"])
    assert ladj.shape == (x.shape[0],) and not ladj.any()


def test_rotation(device):
    t = RotationTransform(dev_t(U["rot_A"], device))
    assert rel_err(cpu(t.R), U["rot_R64"]) < 1e-5
    x = dev_t(U["perm_x"], device)
    assert rel_err(cpu(t(x)), U["rot_y64"]) < 1e-5
    assert rel_err(cpu(t.inv(x)), U["rot_xinv64"]) < 1e-5


def test_diag_normal(device):
    d = zuko.distributions.DiagNormal(dev_t(U["dn_loc"], device), dev_t(U["dn_scale"], device))
    assert rel_err(cpu(d.log_prob(dev_t(U["dn_z"], device))), U["dn_lp64"]) < 2e-6


# --------------------------------------------------------------------------- #
# flows: golden parity (BASELINE configs at reduced batch + the reference's test shapes)
# --------------------------------------------------------------------------- #


def _ctx(g, device, n=None):
    c = g.get("c")
    if c is None:
        return None
    if n is not None and c.ndim > 1:
        c = c[:n]
    return dev_t(c, device)


@pytest.mark.parametrize("name", SMALL_CASES + BIG_CASES)
def test_flow_log_prob_golden(device, math_mode, name):
    g = load(f"flow_{name}")
    flow = build_flow(name, g)
    if "w_scale" in g:
        _set_gemm(flow, "fp32")  # stress set: arbitrated with exact-order fp32 GEMMs
    flow = flow.to(device)
    with torch.no_grad():
        lp = flow(_ctx(g, device)).log_prob(dev_t(g["x"], device))
    assert lp.shape == (g["x"].shape[0],)
    assert_log_prob_parity(cpu(lp), g, rtol=1e-5)


def _set_gemm(flow, mode):
    for t in flow.transform.transforms:
        if hasattr(t, "hyper"):
            t.hyper.gemm_mode = mode


@pytest.mark.parametrize("gemm", ["fp32", "auto"])
@pytest.mark.parametrize("name", SMALL_CASES + BIG_CASES)
def test_flow_forward_and_inverse_golden(device, name, gemm):
    """Per-element z / x and per-sample ladj.  With fp32 conditioner GEMMs the bar is 1e-5
    relative per element; with the tensor-core split-bf16 GEMMs (auto) the operands carry 16
    mantissa bits, so individual elements of z are held to 4e-5 while ladj / log_prob keep
    the 1e-5 bar on the log-density scale (BASELINE.json states the bar on log_prob)."""
    g = load(f"flow_{name}")
    flow = build_flow(name, g)
    _set_gemm(flow, gemm)
    flow = flow.to(device)
    rt = 1e-5 if gemm == "fp32" else 4e-5
    if "w_scale" in g and gemm == "auto":
        pytest.skip("stress set (weights x3) is arbitrated with fp32 GEMMs; see test_stress_set_report")
    t = flow(_ctx(g, device)).transform
    z, ladj = t.call_and_ladj(dev_t(g["x"], device))
    dev32 = np.abs(g["z32"].astype(np.float64) - g["z64"]) if "z32" in g else 0.0
    ez = np.abs(cpu(z) - g["z64"])
    if "w_scale" in g:  # stress set: distributional criterion (see cases.assert_log_prob_parity)
        assert ez.max() <= max(3 * np.max(dev32), 1e-4) and np.median(ez) <= max(3 * np.median(dev32), 1e-6)
    else:
        assert np.all(ez <= np.maximum(rt * np.maximum(np.abs(g["z64"]), 1.0), 3 * dev32)), ez.max()
    devl = np.abs(g["ladj32"].astype(np.float64) - g["ladj64"]) if "ladj32" in g else 0.0
    el = np.abs(cpu(ladj) - g["ladj64"])
    # ladj is a term of log_prob: its error budget is 1e-5 of the log-density magnitude
    lp_scale = np.maximum(np.abs(g["log_prob64"]), 1.0)
    if "w_scale" in g:
        assert el.max() <= max(3 * np.max(devl), 1e-4 * lp_scale.max())
    else:
        assert np.all(el <= np.maximum(1e-5 * lp_scale, 3 * devl)), el.max()
    if "zin" in g:
        n = g["zin"].shape[0]
        ti = flow(_ctx(g, device, n)).transform
        xi = ti.inv(dev_t(g["zin"], device))
        devx = np.abs(g["xinv32"].astype(np.float64) - g["xinv64"]) if "xinv32" in g else 0.0
        ex = np.abs(cpu(xi) - g["xinv64"])
        # |x_ours - x_ref| <= 1e-5 max(1, |x|) on the named configs (x3 the
        # reference's own fp32 deviation where that is larger)
        if "w_scale" in g:
            assert ex.max() <= max(3 * np.max(devx), 1e-3)
        else:
            assert np.all(ex <= np.maximum(2 * rt * np.maximum(np.abs(g["xinv64"]), 1.0), 3 * devx)), ex.max()
        # the reference's own property: t(t.inv(z)) ~ z, atol 1e-4 (tests/test_flows.py:57-61)
        # (the stress set's splines are too sharp for a 1e-4 fp32 round trip: SURVEY §7.4-1b)
        zin = dev_t(g["zin"], device)
        back = ti(xi)
        if name.startswith("ncsf"):  # circular flow: a bijection of [-pi, pi[ only (flows/spline.py:78-80)
            rows = (zin.abs() < np.pi - 1e-3).all(-1)
            zin, back = zin[rows], back[rows]
        assert torch.allclose(back, zin, atol=1e-2 if "w_scale" in g else 1e-4)


def test_stress_set_report(device):
    """Weights x3 (very sharp splines), inputs x3: error distribution of both GEMM modes
    against fp64, next to the reference's own fp32 deviation.  The tensor-core mode is held
    to 2e-3 relative max / 2e-5 median here (documented in DESIGN.md); fp32 mode to the
    distributional criterion in test_flow_log_prob_golden."""
    g = load("flow_nsf6_stress")
    ref64 = g["log_prob64"]
    ref_err = np.abs(g["log_prob32"].astype(np.float64) - ref64) / np.abs(ref64)
    for gemm in ("fp32", "auto"):
        flow = build_flow("nsf6_stress", g)
        _set_gemm(flow, gemm)
        flow = flow.to(device)
        lp = cpu(flow(_ctx(g, device)).log_prob(dev_t(g["x"], device)))
        err = np.abs(lp - ref64) / np.abs(ref64)
        print(f"stress[{gemm}]: max {err.max():.2e} p99 {np.percentile(err, 99):.2e} median {np.median(err):.2e}"
              f" | reference fp32: max {ref_err.max():.2e} p99 {np.percentile(ref_err, 99):.2e} median {np.median(ref_err):.2e}")
        if gemm == "auto":
            assert err.max() < 2e-3 and np.median(err) < 2e-5


def test_per_layer_path_equals_fused_path(device):
    """ComposedTransform's member-by-member path (zk_layer_*) and the one-call flow path
    (zk_flow_*) must agree bit for bit: same kernels, same order."""
    g = load("flow_composed")
    flow = build_flow("composed", g).to(device)
    c, x = _ctx(g, device), dev_t(g["x"], device)
    dist = flow(c)
    z_f, l_f = dist.transform.call_and_ladj(x)
    acc, cur = 0, x
    for t in dist.transform.transforms:
        cur, l = t.call_and_ladj(cur)
        acc = acc + (l.sum(-1) if l.dim() > 1 else l)
    assert torch.equal(cur, z_f)
    assert torch.allclose(acc, l_f, rtol=0, atol=2e-6)
    lp = dist.log_prob(x)
    lp_manual = dist.base.log_prob(z_f) + l_f
    assert torch.allclose(lp, lp_manual, rtol=1e-6, atol=1e-5)


def test_broadcast_context_and_shapes(device):
    """tests/test_flows.py:18-21,41-43: context (5,) broadcast over x (256, 3); sample shapes."""
    g = load("flow_nsf35_row")
    flow = build_flow("nsf35_row", g).to(device)
    c_row = dev_t(g["c"], device)
    x = dev_t(g["x"], device)
    lp_row = flow(c_row).log_prob(x)
    lp_full = flow(c_row.expand(x.shape[0], -1)).log_prob(x)
    assert lp_row.shape == (256,) and torch.equal(lp_row, lp_full)
    # leading dims: (4, 64, 3) with context (64, 5) -> (4, 64)
    c2 = torch.randn(64, 5, device=device)
    lp3 = flow(c2).log_prob(x.reshape(4, 64, 3))
    assert lp3.shape == (4, 64)
    assert torch.equal(lp3[1], flow(c2).log_prob(x.reshape(4, 64, 3)[1]))
    # non-contiguous input
    xt = x.t().contiguous().t()
    assert not xt.is_contiguous() and torch.equal(flow(c_row).log_prob(xt), lp_row)
    s = flow(c_row).sample((32,))
    assert s.shape == (32, 3) and torch.isfinite(s).all()
    s2 = flow(c2).sample((7,))
    assert s2.shape == (7, 64, 3)
    assert flow(c_row).batch_shape == () and flow(c2).batch_shape == (64,)
    assert flow(c_row).event_shape == (3,)


@pytest.mark.parametrize("B", [0, 1, 255, 257, 1000])
def test_ragged_batches(device, B):
    g = load("flow_maf35_batch")
    flow = build_flow("maf35_batch", g).to(device)
    gen = torch.Generator().manual_seed(B)
    x, c = torch.randn(B, 3, generator=gen), torch.randn(B, 5, generator=gen)
    lp = flow(c.to(device)).log_prob(x.to(device))
    assert lp.shape == (B,)
    if B:
        ref = O.flowspec_from_module(build_flow("maf35_batch", g)).log_prob(x.numpy(), c.numpy())
        assert rel_err(cpu(lp), ref) < 1e-5


def test_rsample_and_log_prob(device):
    """distributions.py:129-138: the returned log-density equals log_prob of the sample."""
    g = load("flow_nsf35_row")
    flow = build_flow("nsf35_row", g).to(device)
    dist = flow(dev_t(g["c"], device))
    torch.manual_seed(0)
    x, lp = dist.rsample_and_log_prob((512,))
    assert x.shape == (512, 3) and lp.shape == (512,)
    # the inverse sweep accumulates the ladj itself (fp32 FMA conditioner of ar_inverse.cu); log_prob(x) runs the
    # forward kernels: two arithmetics, each within 1e-5 of the oracle (test_gpu_inverse.py) -> 2e-5 of each other
    assert rel_err(cpu(lp), cpu(dist.log_prob(x))) < 2e-5


def test_chunked_workspace_matches_single_chunk(device):
    """A small workspace forces the flow calls to process the batch in row chunks; the
    result must not change."""
    g = load("flow_cfg2_nsf")
    flow = build_flow("cfg2_nsf", g).to(device)
    gen = torch.Generator().manual_seed(5)
    x, c = torch.randn(20000, 16, generator=gen).to(device), torch.randn(20000, 8, generator=gen).to(device)
    lp_a, tot_a = flow(c).log_prob_and_sum(x)
    old = E.Workspace.max_bytes
    try:
        E.Workspace.clear()
        E.Workspace.max_bytes = 1 << 20
        lp_b, tot_b = flow(c).log_prob_and_sum(x)
    finally:
        E.Workspace.max_bytes = old
        E.Workspace.clear()
    assert torch.equal(lp_a, lp_b)
    assert abs(tot_a.item() - lp_a.double().sum().item()) < 1e-6 * abs(tot_a.item())
    assert tot_a.item() == tot_b.item()  # fixed-order reduction => bit-identical


def test_host_buffer_entry_point(device):
    """zk_flow_log_prob_host: pinned host in, host out, copies pipelined inside the call."""
    g = load("flow_cfg2_nsf")
    flow = build_flow("cfg2_nsf", g).to(device)
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(50000, 16, generator=gen).pin_memory()
    c = torch.randn(50000, 8, generator=gen).pin_memory()
    dist = flow(c.to(device))
    ref = dist.log_prob(x.to(device)).cpu()
    out, total = dist._flow_call()[0].log_prob_host(x, c, device)
    assert torch.equal(out, ref)
    assert abs(total - ref.double().sum().item()) < 1e-6 * abs(total)


def test_host_buffer_entry_point_growing_chunks(device):
    """Past four waves the host pipeline runs growing chunks (1, 2, 3, 5, ... waves; api.cu:host_chunk_plan):
    a ragged 2^18 + 777-row batch, with the preferred workspace and with one that caps the chunks, equals the
    device-resident call bit for bit; the sum is the same fixed-order reduction of per-chunk sums to 1e-9."""
    import ctypes

    g = load("flow_cfg2_nsf")
    flow = build_flow("cfg2_nsf", g).to(device)
    B = (1 << 18) + 777
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(B, 16, generator=gen).pin_memory()
    c = torch.randn(B, 8, generator=gen).pin_memory()
    dist = flow(c.to(device))
    ref = dist.log_prob(x.to(device)).cpu()
    fc = dist._flow_call()[0]
    out, total = fc.log_prob_host(x, c, device)
    assert torch.equal(out, ref)
    assert abs(total - ref.double().sum().item()) < 1e-9 * abs(total)
    # a workspace far below the preferred size: chunks capped, same answer
    L = E.lib()
    small = L.zk_flow_host_workspace_bytes(ctypes.byref(fc.desc), 60000)
    assert small < L.zk_flow_host_workspace_bytes(ctypes.byref(fc.desc), B)
    ws = torch.empty(small, dtype=torch.uint8, device=device)
    out2 = torch.empty(B, dtype=torch.float32).pin_memory()
    tot2 = ctypes.c_double(0.0)
    with torch.cuda.device(device):
        E.check(L.zk_flow_log_prob_host(ctypes.byref(fc.desc), x.data_ptr(), 16, c.data_ptr(), 8, B, out2.data_ptr(),
                                        ctypes.byref(tot2), ws.data_ptr(), ws.numel(), E.stream_ptr(device)))
    assert torch.equal(out2, ref)
    assert abs(tot2.value - total) < 1e-9 * abs(total)


def test_permutations_fold_into_autoregressive_layers(device):
    """A classic MAF stack with explicit PermutationTransform members (zuko/transforms.py:1193-1214 between
    flows/autoregressive.py layers): when no gradient is needed, the permutations are folded into re-indexed
    layer handles (ComposedTransform._folded) — fewer launches, same bijection: log_prob / z / inverse against the
    fp64 oracle and against the member-by-member engine call that the autograd path takes."""
    from functools import partial

    from zuko_b200.flows import MaskedAutoregressiveTransform
    from zuko_b200.lazy import Flow, UnconditionalDistribution, UnconditionalTransform
    from zuko_b200.transforms import MonotonicRQSTransform

    D, C = 6, 2
    torch.manual_seed(31)
    layers = [
        MaskedAutoregressiveTransform(D, C, hidden_features=[64, 64]),
        UnconditionalTransform(PermutationTransform, torch.arange(D).flip(0), buffer=True),
        MaskedAutoregressiveTransform(D, C, univariate=partial(MonotonicRQSTransform, slope=1e-3),
                                      shapes=[(8,), (8,), (7,)], hidden_features=[64, 64]),  # fmt: skip
        UnconditionalTransform(SoftclipTransform, bound=9.0),
        UnconditionalTransform(PermutationTransform, torch.randperm(D), buffer=True),
        MaskedAutoregressiveTransform(D, C, hidden_features=[32], order=torch.randperm(D)),
        UnconditionalTransform(PermutationTransform, torch.randperm(D), buffer=True),
    ]
    base = UnconditionalDistribution(zuko.distributions.DiagNormal, torch.linspace(-0.5, 0.5, D), torch.linspace(0.7, 1.6, D), buffer=True)
    flow_cpu = Flow(layers, base).eval()
    spec = O.flowspec_from_module(flow_cpu)
    gen = torch.Generator().manual_seed(5)
    x, c = torch.randn(3000, D, generator=gen), torch.randn(3000, C, generator=gen)
    ref_lp = spec.log_prob(x.numpy(), c.numpy())
    flow = flow_cpu.to(device)
    xd, cd = x.to(device), c.to(device)
    L = E.lib()
    with torch.no_grad():
        dist = flow(cd)
        fc = dist._flow_call()[0]
        assert fc.folded is not None and len(fc.folded._handles) < len(fc._handles) == 7
        lp = dist.log_prob(xd)  # packs the re-indexed handles
        n0 = L.zk_launch_count()
        lp = dist.log_prob(xd)
        n_folded = L.zk_launch_count() - n0
        z, ladj = dist.transform.call_and_ladj(xd)
        x_back = dist.transform.inv(z)
    lp_reg = flow(cd).log_prob(xd)  # parameters require grad: the member-by-member call with the autograd seam
    assert lp_reg.requires_grad
    with torch.no_grad():
        prev_best = type(fc).best
        try:  # the unfolded call, forward-only, for the launch count
            type(fc).best = lambda self, x, c: self
            flow(cd).log_prob(xd)
            n0 = L.zk_launch_count()
            flow(cd).log_prob(xd)
            n_plain = L.zk_launch_count() - n0
            z_reg, ladj_reg = flow(cd).transform.call_and_ladj(xd)
        finally:
            type(fc).best = prev_best
    assert n_folded < n_plain
    assert rel_err(cpu(lp), ref_lp) < 1e-5 and rel_err(cpu(lp_reg.detach()), ref_lp) < 1e-5
    assert torch.allclose(lp, lp_reg.detach(), rtol=1e-5, atol=2e-5)
    assert torch.allclose(z, z_reg, rtol=1e-5, atol=1e-5) and torch.allclose(ladj, ladj_reg, rtol=1e-5, atol=2e-5)
    assert (x_back - xd).abs().max().item() < 1e-4
    (-lp_reg.mean()).backward()  # the autograd path is untouched by the fold
    assert all(p.grad is not None for p in flow.parameters())


def test_circular_shift_and_box_uniform(device):
    """CircularShiftTransform (transforms.py:319-351) and BoxUniform.log_prob (distributions.py:366-396)
    against the reference's golden vectors: the shift is bit-exact in fp32."""
    from zuko_b200.distributions import BoxUniform
    from zuko_b200.transforms import CircularShiftTransform

    UN = load("units_ncsf")
    x = dev_t(UN["circ_x"], device)
    for tag, b in (("circ1", 1.0), ("circpi", float(np.pi))):
        t = CircularShiftTransform(bound=b)
        y, ladj = t.call_and_ladj(x)
        assert np.array_equal(y.cpu().numpy(), UN[f"{tag}_y32"]) and float(ladj.abs().max()) == 0.0
        assert np.array_equal(t.inv(x).cpu().numpy(), UN[f"{tag}_y32"])
    box = BoxUniform(dev_t(UN["box_lower"], device), dev_t(UN["box_upper"], device))
    lp = cpu(box.log_prob(dev_t(UN["box_z"], device)))
    ref = UN["box_lp64"]
    assert np.array_equal(np.isinf(lp), np.isinf(ref))
    assert rel_err(lp[np.isfinite(ref)], ref[np.isfinite(ref)]) < 1e-6
    s = box.sample((1000,))
    assert s.shape == (1000, 3) and bool(((s >= box.lower) & (s < box.upper)).all())


def test_parameter_update_repacks(device):
    """The packed weights are a cache keyed on (data_ptr, version): an in-place update (an
    optimizer step) must be seen by the next call."""
    g = load("flow_maf35_batch")
    flow = build_flow("maf35_batch", g).to(device)
    x, c = dev_t(g["x"], device), dev_t(g["c"], device)
    lp0 = flow(c).log_prob(x)
    with torch.no_grad():
        flow.transform.transforms[0].hyper[4].bias.add_(0.25)
    lp1 = flow(c).log_prob(x)
    assert not torch.allclose(lp0, lp1)
    cpu_flow = build_flow("maf35_batch", g)
    with torch.no_grad():
        cpu_flow.transform.transforms[0].hyper[4].bias.add_(0.25)
    ref = O.flowspec_from_module(cpu_flow).log_prob(g["x"], g["c"])
    assert rel_err(cpu(lp1), ref) < 1e-5


def test_errors(device):
    g = load("flow_maf35_batch")
    flow = build_flow("maf35_batch", g).to(device)
    x, c = dev_t(g["x"], device), dev_t(g["c"], device)
    with pytest.raises(TypeError):
        flow(c).log_prob(x.double())
    xs = flow(c).transform.inv(x.clone().requires_grad_())  # the inverse direction is differentiable too
    assert xs.requires_grad
    with pytest.raises(ValueError):
        flow(c).log_prob(x[:, :2])
    with pytest.raises(E.EngineError):
        flow(c).log_prob(x.cpu())
    with pytest.raises((ValueError, E.EngineError)):
        flow(None).log_prob(x)  # the flow needs a context


def test_masked_mlp_jacobian_sparsity(device):
    """tests/test_nn.py:39-60: outputs must not react to inputs their adjacency row forbids."""
    torch.manual_seed(0)
    adjacency = torch.rand(6, 5) < 0.4
    adjacency[:, 0] = True
    net = zuko.nn.MaskedMLP(adjacency, [32, 48]).to(device)
    x = torch.randn(5, device=device)
    base = net(x)
    for j in range(5):
        xp = x.clone()
        xp[j] += 1.0
        moved = (net(xp) - base).abs() > 0
        assert not (moved.cpu() & ~adjacency[:, j]).any(), j


# --------------------------------------------------------------------------- #
# BASELINE sizes: size-independent properties
# --------------------------------------------------------------------------- #


def test_cfg2_full_batch_properties(device):
    """B = 2^20 (BASELINE config 2).  (i) the first 4096 rows match the oracle at 1e-5;
    (ii) the batch is 16 copies of one 2^16-row block, so every block must reproduce the
    first bit for bit (tiling / chunking independence); (iii) the fused sum equals the
    host double sum of the per-sample values."""
    g = load("flow_cfg2_nsf")
    cpu_flow = build_flow("cfg2_nsf", g)
    flow = build_flow("cfg2_nsf", g).to(device)
    gen = torch.Generator().manual_seed(1234)
    xb, cb = torch.randn(1 << 16, 16, generator=gen), torch.randn(1 << 16, 8, generator=gen)
    x, c = xb.repeat(16, 1).to(device), cb.repeat(16, 1).to(device)
    lp, total = flow(c).log_prob_and_sum(x)
    assert lp.shape == (1 << 20,)
    ref = O.flowspec_from_module(cpu_flow).log_prob(xb[:4096].numpy(), cb[:4096].numpy())
    assert rel_err(cpu(lp[:4096]), ref) < 1e-5
    blocks = lp.reshape(16, 1 << 16)
    assert all(torch.equal(blocks[0], blocks[i]) for i in range(1, 16))
    assert abs(total.item() - lp.double().sum().item()) <= 1e-9 * abs(total.item())


def _full_batch_properties(device, name, D, C, block_rows, copies, oracle_rows=2048):
    g = load(f"flow_{name}")
    cpu_flow = build_flow(name, g)
    flow = build_flow(name, g).to(device)
    gen = torch.Generator().manual_seed(1234)
    xb = torch.randn(block_rows, D, generator=gen)
    cb = torch.randn(block_rows, C, generator=gen) if C else None
    x = xb.repeat(copies, 1).to(device)
    c = None if cb is None else cb.repeat(copies, 1).to(device)
    n0 = E.lib().zk_launch_count()
    lp, total = flow(c).log_prob_and_sum(x)
    assert lp.shape == (block_rows * copies,)
    ref = O.flowspec_from_module(cpu_flow).log_prob(xb[:oracle_rows].numpy(), None if cb is None else cb[:oracle_rows].numpy())
    assert rel_err(cpu(lp[:oracle_rows]), ref) < 1e-5
    blocks = lp.reshape(copies, block_rows)
    assert all(torch.equal(blocks[0], blocks[i]) for i in range(1, copies))
    assert abs(total.item() - lp.double().sum().item()) <= 1e-9 * abs(total.item())
    return flow, x, c


def test_cfg3_full_batch_properties(device):
    """BASELINE config 3 at its full size, B = 2^20 (MAF(32, T8, [512]^4): the wide fused kernel):
    oracle parity on the first rows, block-for-block bit reproducibility across the batch (tiling /
    chunking independence), and the fused sum against the host double sum."""
    _full_batch_properties(device, "cfg3_maf", 32, 0, 1 << 16, 16)


def test_cfg5_full_shard_properties(device):
    """BASELINE config 5 at one GPU's share of the 2^24-row batch, B = 2^21
    (NSF(64, 16, T8, K16, [512]^3): wide fused kernel, 47 parameters per dim)."""
    _full_batch_properties(device, "cfg5_nsf", 64, 16, 1 << 17, 16, oracle_rows=1024)


def test_cfg4_round_trip_2_20(device):
    """BASELINE config 4 at its full size: t(t.inv(z)) ~ z for N = 2^20 draws (the reference's own
    property, atol 1e-4, tests/test_flows.py:57-61), and the first rows against the oracle's inverse."""
    g = load("flow_cfg4_nsf")
    cpu_flow = build_flow("cfg4_nsf", g)
    flow = build_flow("cfg4_nsf", g).to(device)
    gen = torch.Generator().manual_seed(77)
    z = torch.randn(1 << 20, 64, generator=gen)
    t = flow(None).transform
    zd = z.to(device)
    x = t.inv(zd)
    assert torch.isfinite(x).all()
    back = t(x)
    assert torch.allclose(back, zd, atol=1e-4)
    ref = O.flowspec_from_module(cpu_flow).inverse(z[:512].numpy(), None)
    assert rel_err(cpu(x[:512]), ref) < 1e-5


def test_cfg2_round_trip_2_18(device):
    """inv(t(x)) ~ x at 2^18 rows (the reference's atol 1e-4, tests/test_flows.py:57-61)."""
    g = load("flow_cfg2_nsf")
    flow = build_flow("cfg2_nsf", g).to(device)
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(1 << 18, 16, generator=gen).to(device)
    c = torch.randn(1 << 18, 8, generator=gen).to(device)
    t = flow(c).transform
    assert torch.allclose(t.inv(t(x)), x, atol=1e-4)
