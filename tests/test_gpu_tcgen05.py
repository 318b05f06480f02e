"""GPU tests of the tcgen05 (tensor-core) conditioner path against the fp64 oracle and the
fp32 CUDA-core path."""

import numpy as np
import pytest
import torch

import zuko_b200 as zuko
from cases import assert_log_prob_parity, build_flow, load, rel_err
from oracle import oracle as O
from zuko_b200 import _engine as E

pytestmark = pytest.mark.gpu


def _oracle_mlp(net, x):
    cond = O.conditioner_from_module(net)
    return cond(x.numpy().astype(np.float64), None, np.float64)


def _masked(adj_shape, hidden, seed=0):
    torch.manual_seed(seed)
    out_f, in_f = adj_shape
    adjacency = torch.rand(out_f, in_f) < 0.6
    adjacency[:, 0] = True
    return zuko.nn.MaskedMLP(adjacency, hidden)


SHAPES = [
    ((368, 24), [256, 256, 256]),  # BASELINE config 2 conditioner
    ((69, 8), [64, 64]),           # N not a multiple of 16, K padded 8 -> 64
    ((64, 32), [512, 512]),        # two N chunks per hidden layer
    ((6, 100), [96]),              # hidden width 96 -> padded to 128 for the next K
    ((300, 64), []),               # single linear layer
]


@pytest.mark.parametrize("shape,hidden", SHAPES)
@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000, 40000])
def test_tcgen05_mlp_vs_oracle(device, shape, hidden, B):
    net = _masked(shape, hidden)
    net.gemm_mode = "bf16x3"
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, shape[1], generator=g)
    ref = _oracle_mlp(net, x)
    out = net.to(device)(x.to(device))
    assert E.lib().zk_mlp_gemm_mode(net._handle()) == E.ZK_GEMM_BF16X3
    out = out.cpu().numpy().astype(np.float64)
    scale = np.abs(ref).max()
    err = np.abs(out - ref).max() / scale
    # split-bf16 (3 MMAs, fp32 accumulate): ~2^-16 per product term; bound 3e-5 of the output scale
    assert err < 3e-5, err
    # and the fp32 CUDA-core path agrees with both
    net32 = _masked(shape, hidden)
    net32.gemm_mode = "fp32"
    out32 = net32.to(device)(x.to(device)).cpu().numpy().astype(np.float64)
    assert np.abs(out32 - ref).max() / scale < 2e-6


def test_bf16x1_is_available_but_inexact(device):
    net = _masked((368, 24), [256, 256, 256])
    net.gemm_mode = "bf16x1"
    x = torch.randn(4096, 24, generator=torch.Generator().manual_seed(1))
    ref = _oracle_mlp(net, x)
    out = net.to(device)(x.to(device)).cpu().numpy().astype(np.float64)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert 1e-4 < err < 3e-2, err  # single bf16 MMA: ~2^-8 per operand


@pytest.mark.parametrize("name", ["cfg2_nsf", "cfg3_maf", "cfg4_nsf", "cfg5_nsf", "nsf35_row", "nice35"])
def test_flow_parity_on_tensor_cores(device, name):
    """The BASELINE configs with the conditioner forced onto tcgen05 (bf16x3): log_prob within
    1e-5 relative of the reference."""
    g = load(f"flow_{name}")
    flow = build_flow(name, g)
    for t in flow.transform.transforms:
        t.hyper.gemm_mode = "bf16x3"
    flow = flow.to(device)
    c = None if "c" not in g else torch.from_numpy(g["c"]).to(device)
    lp = flow(c).log_prob(torch.from_numpy(g["x"]).to(device))
    assert E.lib().zk_mlp_gemm_mode(flow.transform.transforms[0].hyper._handle()) == E.ZK_GEMM_BF16X3
    assert_log_prob_parity(lp.cpu().numpy(), g, rtol=1e-5)


def test_auto_mode_picks_tensor_cores_for_wide_layers(device):
    wide = _masked((368, 24), [256, 256]).to(device)
    tiny = _masked((8, 4), [32, 32]).to(device)
    assert E.lib().zk_mlp_gemm_mode(wide._handle()) == E.ZK_GEMM_BF16X3
    assert E.lib().zk_mlp_gemm_mode(tiny._handle()) == E.ZK_GEMM_FP32
