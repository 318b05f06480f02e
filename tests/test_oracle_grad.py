"""CPU tests: the gradient oracle (oracle/oracle_grad.py) against gradients that torch.autograd
produced on the unmodified reference in fp64 (tests/golden/make_golden_grad.py).  This PINS the
hand-written reverse-mode restatement the GPU backward pass is checked against."""

import numpy as np
import pytest

from cases import INV_GRAD_CASES, inv_grad_inputs, GRAD_CASES_FULL, GRAD_CASES_SAMPLED, assert_param_grads, build_flow, grad_inputs, load, oracle_named_grads
from oracle import oracle as O
from oracle import oracle_grad as OG

GU = load("grad_units")
U = load("units")
T = dict(rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("tag", ["s01", "s1", "s3"])
def test_rqs_backward_batched(tag):
    x = GU[f"rqs_{tag}_x"]
    gx, gphi = OG.rqs_backward(x, U[f"rqs_{tag}_phi"], GU[f"rqs_{tag}_gy"], GU[f"rqs_{tag}_gl"], 8)
    # x exactly ON the last knot (+5): whether it counts as inside depends on the last ulp of
    # cumsum(softmax) — not a property of the derivative; skip that element
    ok = np.abs(x) != 5.0
    np.testing.assert_allclose(gx[ok], GU[f"rqs_{tag}_gx"][ok], **T)
    np.testing.assert_allclose(gphi[ok], GU[f"rqs_{tag}_gphi"][ok], **T)


@pytest.mark.parametrize("K", [16, 5])
def test_rqs_backward_shared_table(K):
    gx, gphi = OG.rqs_backward(U[f"rqs_shared{K}_x"], U[f"rqs_shared{K}_phi"], GU[f"rqs_shared{K}_gy"], GU[f"rqs_shared{K}_gl"], K)
    np.testing.assert_allclose(gx, GU[f"rqs_shared{K}_gx"], **T)
    np.testing.assert_allclose(gphi, GU[f"rqs_shared{K}_gphi"], **T)


def test_affine_backward():
    gx, gphi = OG.affine_backward(U["affine_x"], U["affine_phi"], GU["affine_gy"], GU["affine_gl"])
    np.testing.assert_allclose(gx, GU["affine_gx"], **T)
    np.testing.assert_allclose(gphi, GU["affine_gphi"], **T)


@pytest.mark.parametrize("bound", [1, 11])
def test_softclip_backward(bound):
    gx = OG.softclip_backward(U["softclip_x"], GU["softclip_gy"], GU["softclip_gl"], float(bound))
    np.testing.assert_allclose(gx, GU[f"softclip{bound}_gx"], **T)


@pytest.mark.parametrize("name", GRAD_CASES_FULL + GRAD_CASES_SAMPLED)
def test_flow_gradients(name):
    gg, x, c = grad_inputs(name)
    flow = build_flow(name)
    spec = O.flowspec_from_module(flow)
    # d/d(x, c, theta) of sum_b g_b log_prob_b
    gx, gc, lgs = OG.flow_backward(spec, x, c, g_log_prob=gg["g"])
    np.testing.assert_allclose(gx, gg["lp/gx"], rtol=1e-7, atol=1e-7)
    if c is not None:
        np.testing.assert_allclose(gc, gg["lp/gc"], rtol=1e-7, atol=1e-7)
    assert_param_grads(oracle_named_grads(flow, lgs), gg, "lp/", 1e-8, name)
    # d/d(x, c, theta) of <gz, z> + <gl, ladj>
    gx, gc, lgs = OG.flow_backward(spec, x, c, g_z=gg["gz"], g_ladj=gg["gl"])
    np.testing.assert_allclose(gx, gg["tr/gx"], rtol=1e-7, atol=1e-7)
    if c is not None:
        np.testing.assert_allclose(gc, gg["tr/gc"], rtol=1e-7, atol=1e-7)
    assert_param_grads(oracle_named_grads(flow, lgs), gg, "tr/", 1e-8, name)


@pytest.mark.parametrize("name", INV_GRAD_CASES)
def test_inverse_direction_gradients(name):
    """Implicit differentiation of x = transform.inv(z) (and of rsample_and_log_prob's log-density)
    against torch.autograd back-propagating through the reference's inverse sweeps."""
    gg, z, c = inv_grad_inputs(name)
    flow = build_flow(name)
    spec = O.flowspec_from_module(flow)
    np.testing.assert_allclose(spec.inverse(gg["z"], c), gg["x64"], rtol=1e-8, atol=1e-8)
    for prefix, kw in (("inv/", dict(g_x=gg["w"])), ("invlp/", dict(g_x=gg["w"], g_log_prob=gg["wl"]))):
        gz, gc, lgs = OG.flow_inverse_backward(spec, gg["z"], c, **kw)
        np.testing.assert_allclose(gz, gg[prefix + "gx"], rtol=1e-7, atol=1e-7)
        if c is not None:
            np.testing.assert_allclose(gc, gg[prefix + "gc"], rtol=1e-7, atol=1e-7)
        assert_param_grads(oracle_named_grads(flow, lgs), gg, prefix, 1e-8, name)
