"""CPU tests of the host side: bit-exact mask / order construction, constructor parity
with the reference (state-dict keys + CRC32 of every tensor), the C-ABI library
(loads, exports every declared symbol), and the no-fallback behaviour."""

import ctypes
import io
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from cases import FLOW_CASES, build_flow, crc, load
from zuko_b200 import _engine as E
from zuko_b200.flows import MAF, NSF, MaskedAutoregressiveTransform
from zuko_b200.nn import MaskedMLP, masked_mlp_masks
from zuko_b200.transforms import MonotonicRQSTransform

ROOT = Path(__file__).resolve().parent.parent
M = load("masks")


def _mat_specs():
    adjacency = torch.from_numpy(M["adjacency"])
    adj_ctx = torch.from_numpy(M["adjacency_ctx"])
    rqs8 = dict(univariate=MonotonicRQSTransform, shapes=[(8,), (8,), (7,)])
    rqs16 = dict(univariate=MonotonicRQSTransform, shapes=[(16,), (16,), (15,)])
    return {
        "maf4": dict(features=4, context=0, hidden_features=[32, 32]),
        "nsf16c8": dict(features=16, context=8, hidden_features=[256] * 3, **rqs8),
        "passes2": dict(features=5, context=7, passes=2, hidden_features=[16, 24]),
        "order": dict(features=5, context=0, order=[3, 0, 4, 1, 2], hidden_features=[17]),
        "rev64": dict(features=64, context=0, order=list(range(63, -1, -1)), hidden_features=[64, 64], **rqs16),
        "adjacency": dict(features=5, context=0, adjacency=adjacency, hidden_features=[12, 12]),
        "adjacency_ctx": dict(features=5, context=3, adjacency=adj_ctx, hidden_features=[12]),
    }


@pytest.mark.parametrize("name", list(_mat_specs()))
def test_masks_bit_exact(name):
    """Integer / boolean work must be bit-exact: masks, order classes and passes equal the
    reference's (zuko/nn.py:258-318, flows/autoregressive.py:106-152)."""
    t = MaskedAutoregressiveTransform(**_mat_specs()[name])
    assert t.passes == int(M[f"{name}/passes"])
    if f"{name}/order" in M:
        assert np.array_equal(t.order.numpy(), M[f"{name}/order"])
    else:
        assert t.order is None
    linears = [m for m in t.hyper if hasattr(m, "mask")]
    for i, m in enumerate(linears):
        shape = tuple(M[f"{name}/shape{i}"])
        assert tuple(m.mask.shape) == shape
        ref = np.unpackbits(M[f"{name}/mask{i}"])[: shape[0] * shape[1]].reshape(shape).astype(bool)
        assert np.array_equal(m.mask.numpy(), ref), (name, i)
    assert f"{name}/mask{len(linears)}" not in M


def test_free_standing_masked_mlp():
    adj = torch.from_numpy(M["free/adjacency"])
    masks = masked_mlp_masks(adj, [16, 32])
    for i, m in enumerate(masks):
        assert np.array_equal(m.numpy(), M[f"free/mask{i}"])
    with pytest.raises(ValueError):
        masked_mlp_masks(torch.zeros(3, 3, dtype=torch.bool), [4])


def test_adjacency_validation():
    # reference: tests/test_flows.py:147-218
    with pytest.raises(AssertionError, match="diagonal"):
        MaskedAutoregressiveTransform(3, adjacency=torch.zeros(3, 3, dtype=torch.bool))
    cyc = torch.eye(3, dtype=torch.bool)
    cyc[0, 1] = cyc[1, 2] = cyc[2, 0] = True
    with pytest.raises(AssertionError, match="cycles"):
        MaskedAutoregressiveTransform(3, adjacency=cyc)
    with pytest.raises(AssertionError, match="columns"):
        MaskedAutoregressiveTransform(3, 2, adjacency=torch.eye(3, 4, dtype=torch.bool))


@pytest.mark.parametrize("name", list(FLOW_CASES))
def test_constructor_parity(name):
    """Same seed => same state-dict keys and bit-identical tensors as the reference."""
    g = load(f"flow_{name}")
    seed, ctor = FLOW_CASES[name]
    if seed is not None:
        torch.manual_seed(seed)
    sd = ctor().state_dict()
    assert list(sd.keys()) == list(g["sd_keys"])
    ours = np.array([crc(v) for v in sd.values()], dtype=np.int64)
    bad = [k for k, a, b in zip(sd.keys(), ours, g["sd_crc"]) if a != b]
    assert not bad, f"tensors differ from the reference initialisation: {bad[:5]}"


def test_reference_checkpoint_loads():
    g = load("flow_cfg1_maf")
    flow = MAF(4, 0, transforms=2, hidden_features=[32, 32])
    stored = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
    flow.load_state_dict(stored, strict=True)
    buf = io.BytesIO()
    torch.save(flow, buf)  # tests/test_flows.py:77-91 saves the whole module
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert all(torch.equal(a, b) for a, b in zip(flow.state_dict().values(), again.state_dict().values()))
    assert "MaskedAutoregressiveTransform" in repr(again)


def test_features_one_falls_back_to_elementwise():
    # flows/autoregressive.py:73-86
    t = MaskedAutoregressiveTransform(1, 3)
    assert type(t).__name__ == "ElementWiseTransform"


def test_unsupported_options_raise():
    with pytest.raises(NotImplementedError):
        MaskedMLP(torch.ones(4, 3, dtype=torch.bool), activation=torch.nn.PReLU)
    with pytest.raises(NotImplementedError):
        MaskedMLP(torch.ones(4, 3, dtype=torch.bool), activation=lambda: torch.nn.ELU(alpha=0.5))
    with pytest.raises(NotImplementedError):
        MaskedAutoregressiveTransform(3, univariate=torch.distributions.ExpTransform, shapes=[])


def test_residual_masked_mlp_structure():
    """MaskedMLP(residual=True) builds the reference's blocks (zuko/nn.py:297-309): module tree, per-layer
    activation / residual flags handed to the engine, and the masks keep the Jacobian sparsity of the
    adjacency (the reference's own property test, tests/test_nn.py:39-60, in exact arithmetic)."""
    torch.manual_seed(1)
    adjacency = torch.rand(6, 5) < 0.4
    adjacency[:, 0] = True
    net = MaskedMLP(adjacency, [16, 16, 24], activation=torch.nn.ELU, residual=True)
    names = [type(m).__name__ for m in net]
    assert names == ["MaskedLinear", "Residual", "Residual", "MaskedLinear", "Residual", "MaskedLinear"], names
    acts, res = net._layer_flags()
    assert acts == [0, 2, 0, 2, 0, 0, 2, 0, 0] and res == [0, 0, 1, 0, 1, 0, 0, 1, 0], (acts, res)
    # boolean reachability through the masks = structural Jacobian; residual adds keep a unit's own class
    lins = net._linears()
    reach = torch.eye(5, dtype=torch.bool)  # (unit of the current layer input) x (network input)
    prev_in = None
    for lin, r in zip(lins, res, strict=True):
        cur_in = reach
        out = (lin.mask.to(torch.float64) @ reach.to(torch.float64)) > 0
        if r:
            out = out | prev_in
        prev_in, reach = cur_in, out
    assert not (reach & ~adjacency).any()


# --------------------------------------------------------------------------- #
# the C-ABI library
# --------------------------------------------------------------------------- #


def _declared_symbols():
    header = (ROOT / "include" / "zuko_b200.h").read_text()
    pat = r"^(?:zk_status|int|int64_t|size_t|void|const char\*)\s+(zk_[a-z0-9_]+)\s*\("
    return sorted(set(re.findall(pat, header, flags=re.M)))


def test_library_exports_every_declared_symbol():
    lib = E.lib()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/zuko_b200.h but not exported"
    assert set(E.EXPORTED_SYMBOLS) == set(declared), set(E.EXPORTED_SYMBOLS) ^ set(declared)
    assert lib.zk_version() >= 100


def test_no_cpu_fallback():
    """CPU tensors are refused loudly; nothing silently routes through torch or the oracle."""
    flow = build_flow("nsf35_row")
    x, c = torch.randn(8, 3), torch.randn(5)
    with pytest.raises(E.EngineError, match="no CPU fallback"):
        flow(c).log_prob(x)
    with pytest.raises(E.EngineError, match="no CPU fallback"):
        _ = zuko.transforms.SoftclipTransform()(torch.randn(3))


def test_compute_without_device_is_an_error():
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = E.lib()
    st = lib.zk_softclip_forward(None, 1, 0, 1, ctypes.c_float(1.0), None, 1, None, 0, None)
    assert st in (E.ZK_EINVAL, E.ZK_ECUDA)
    sm = ctypes.c_int()
    assert lib.zk_device_info(ctypes.byref(sm), None, None) == E.ZK_ECUDA


def test_product_does_not_import_oracle():
    for path in (ROOT / "zuko_b200").rglob("*.py"):
        src = path.read_text()
        assert "oracle" not in src.replace("no CPU", ""), f"{path} mentions the oracle"


def test_host_pipeline_chunk_plan():
    """zk_flow_log_prob_host's row chunks (api.cu:host_chunk_plan): they cover the batch exactly, start with one
    wave, are whole waves except the last, never grow by more than 2x (the copy of chunk k has to fit inside the
    compute of chunk k-1), respect the cap the workspace imposes, and small batches keep the eight-equal-chunks plan."""
    import numpy as np

    from zuko_b200 import _engine as E

    wave = 74 * 512

    def plan(B, cap=None):
        out = np.zeros(4096, np.int64)
        n = E.lib().zk_debug_host_chunk_plan(B, wave, cap or B, out.ctypes.data, 4096)
        assert 0 <= n <= 4096
        return out[:n].tolist()

    assert plan(0) == []
    assert plan(1) == [1]
    assert plan(5000) == [4096, 904]
    assert plan(100000) == [12500] * 8
    for B in (1 << 18, 1 << 20, (1 << 20) + 12345, 1 << 21, 1 << 24):
        p = plan(B)
        assert sum(p) == B and p[0] == wave
        assert all(n % wave == 0 for n in p[:-1])
        assert all(b <= 2 * a for a, b in zip(p, p[1:-1]))
        assert p[-1] >= wave // 2 or len(p) == 1
        assert max(p) <= 32 * wave + wave // 2
    p = plan(1 << 20, cap=200000)
    assert sum(p) == 1 << 20 and max(p) <= 200000 and p[0] == wave
    assert sum(plan(1 << 20, cap=1000)) == 1 << 20 and max(plan(1 << 20, cap=1000)) <= 1000


@pytest.mark.parametrize("rqs", [False, True])
def test_folding_a_permutation_into_an_autoregressive_layer(rqs):
    """The algebra behind ComposedTransform._folded, on the CPU with the oracle: for (P x)_j = x[q[j]] and the layer
    T~ re-indexed by flows.autoregressive.reindex_conditioner (+ order~[q] = order),  T(P x) = P T~(x) with equal
    ladj — for the forward and the inverse direction."""
    import copy
    from functools import partial

    import numpy as np

    from oracle import oracle as O
    from zuko_b200.flows import MaskedAutoregressiveTransform
    from zuko_b200.flows.autoregressive import reindex_conditioner
    from zuko_b200.lazy import Flow, UnconditionalDistribution
    from zuko_b200.transforms import MonotonicRQSTransform

    D, C = 6, 2
    torch.manual_seed(17)
    kw = dict(univariate=partial(MonotonicRQSTransform, slope=1e-3), shapes=[(4,), (4,), (3,)]) if rqs else {}
    t = MaskedAutoregressiveTransform(D, C, hidden_features=[16, 16], order=torch.randperm(D), **kw)
    q = torch.randperm(D)
    t2 = copy.deepcopy(t)
    lins, lins2 = t.hyper._linears(), t2.hyper._linears()
    with torch.no_grad():
        for i, (a, b) in enumerate(zip(lins, lins2)):
            w, bias, mask = reindex_conditioner(i, len(lins), a.weight.detach(), a.bias.detach(), a.mask, qi=q, D=D, P=t.total)
            b.weight.copy_(w)
            b.bias.copy_(bias)
            b.mask = mask
        ro = torch.empty_like(t.order)
        ro[q] = t.order
        t2.order = ro
    base = lambda: UnconditionalDistribution(zuko.distributions.DiagNormal, torch.zeros(D), torch.ones(D), buffer=True)  # noqa: E731
    s1, s2 = O.flowspec_from_module(Flow([t], base()).eval()), O.flowspec_from_module(Flow([t2], base()).eval())
    g = np.random.default_rng(0)
    x, c = g.standard_normal((64, D)), g.standard_normal((64, C))
    qn = q.numpy()
    y1, l1 = s1.forward(x[:, qn], c)      # T(P x)
    y2, l2 = s2.forward(x, c)             # T~(x)
    assert np.allclose(y1, y2[:, qn], rtol=1e-12, atol=1e-12) and np.allclose(l1, l2, rtol=1e-12, atol=1e-12)
    # inverse: T^-1(P z) = P T~^-1(z)
    z = g.standard_normal((64, D)) * 0.5
    assert np.allclose(s1.inverse(z[:, qn], c), s2.inverse(z, c)[:, qn], rtol=1e-9, atol=1e-9)


def test_permutation_fold_plan_is_the_same_bijection():
    """transforms.plan_permutation_fold on random member sequences, simulated with numpy: "reindex" members are
    order-dependent maps f_i (so that a wrong re-indexing shows), their re-indexed form is f~(s) = P^-1 f(P s);
    "commute" members act element-wise, "fixed" members are order-dependent and may not be re-indexed.  The planned
    sequence must give the sequential result, and never runs more members than the original."""
    import numpy as np

    from zuko_b200.transforms import plan_permutation_fold

    rng = np.random.default_rng(0)
    D = 7
    n_plans = 0
    for trial in range(300):
        n = int(rng.integers(1, 9))
        kinds = [str(rng.choice(["perm", "perm", "reindex", "reindex", "commute", "fixed"])) for _ in range(n)]
        sigmas = [rng.permutation(D).tolist() if k == "perm" else None for k in kinds]
        mats = [rng.standard_normal((D, D)) for _ in range(n)]  # an order-dependent map per member

        def apply(i, v):
            if kinds[i] == "perm":
                return v[sigmas[i]]
            if kinds[i] == "commute":
                return np.tanh(v) + 0.1 * i
            return np.tanh(mats[i] @ v) + np.arange(D) * 0.01  # "reindex" / "fixed": depends on the feature order

        x = rng.standard_normal(D)
        ref = x.copy()
        for i in range(n):
            ref = apply(i, ref)
        plan = plan_permutation_fold(kinds, sigmas, D)
        if plan is None:
            continue
        n_plans += 1
        assert len(plan) < n
        s = x.copy()
        for act in plan:
            if act[0] == "gather":
                s = s[list(act[1])]
                continue
            _, i, q = act
            assert kinds[i] != "perm" and (q is None or kinds[i] == "reindex")
            if q is None:
                s = apply(i, s)
            else:  # T~(s) = P^-1 T(P s) with (P v)_j = v[q[j]]
                qq = list(q)
                out = np.empty(D)
                out[qq] = apply(i, s[qq])
                s = out
        assert np.allclose(s, ref, rtol=0, atol=1e-12), (kinds, plan)
    assert n_plans > 50
    # two reversals around a layer cancel: no gather at all
    rev = list(range(D))[::-1]
    assert plan_permutation_fold(["perm", "reindex", "perm"], [rev, None, rev], D) == [("member", 1, tuple(rev))]
    assert plan_permutation_fold(["perm", "fixed"], [rev, None], D) is None          # gather + member: nothing saved
    assert plan_permutation_fold(["perm", "reindex"], [[0, 0, 1, 2, 3, 4, 5], None], D) is None  # not a permutation


def test_flow_call_picks_the_folded_sibling_only_without_gradients():
    """_ops.FlowCall.best: the permutation-folded (forward-only) sibling serves a call exactly when nothing has to be
    differentiated — no grad mode, or no input / context / parameter / base tensor that requires grad."""
    from zuko_b200 import _ops

    w = torch.nn.Parameter(torch.zeros(3))
    plain = _ops.FlowCall([1, 2, 3], 4, 2, None, None, sources=[{"weights": [w], "biases": [None]}, {}, {}])
    folded = _ops.FlowCall([1, 3], 4, 2, None, None, sources=None)
    x, c = torch.zeros(5, 4), torch.zeros(5, 2)
    assert plain.best(x, c) is plain  # nothing folded yet
    plain.folded = folded
    with torch.no_grad():
        assert plain.best(x, c) is folded
        assert plain.best(x.clone().requires_grad_(), c) is folded
    assert plain.best(x, c) is plain  # grad mode and a parameter that requires grad
    w.requires_grad_(False)
    assert plain.best(x, c) is folded
    assert plain.best(x.clone().requires_grad_(), c) is plain
    assert plain.best(x, c.clone().requires_grad_()) is plain
    assert folded.best(x, c) is folded and folded.usable(x, c)
