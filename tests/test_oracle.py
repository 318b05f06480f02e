"""CPU tests: the oracle (oracle/) against the golden vectors of the Python reference.

This is what PINS the oracle: every function of the C restatement is compared with
outputs of the unmodified reference (tests/golden/make_golden.py) — tightly in fp64,
and at fp32 round-off in fp32.
"""

import numpy as np
import pytest

from cases import FLOW_CASES, build_flow, load
from oracle import oracle as O

U = load("units")
T64 = dict(rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("tag", ["s01", "s1", "s3"])
def test_rqs_knots(tag):
    X, Y, Dv = O.rqs_knots(U[f"rqs_{tag}_phi"], 8)
    np.testing.assert_allclose(X, U[f"rqs_{tag}_horizontal"], **T64)
    np.testing.assert_allclose(Y, U[f"rqs_{tag}_vertical"], **T64)
    np.testing.assert_allclose(Dv, U[f"rqs_{tag}_derivatives"], **T64)


@pytest.mark.parametrize("tag", ["s01", "s1", "s3"])
def test_rqs_forward_inverse_f64(tag):
    phi, x = U[f"rqs_{tag}_phi"], U[f"rqs_{tag}_x"]
    N, D = x.shape
    y, ladj = O.rqs_forward(x, phi.reshape(N, -1), 8)
    np.testing.assert_allclose(y, U[f"rqs_{tag}_y64"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ladj, U[f"rqs_{tag}_ladj64"], rtol=1e-8, atol=1e-8)
    xi = O.rqs_inverse(U[f"rqs_{tag}_yq"], phi.reshape(N, -1), 8)
    np.testing.assert_allclose(xi, U[f"rqs_{tag}_inv_of_yq64"], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("tag,tol", [("s01", 2e-5), ("s1", 2e-3)])
def test_rqs_forward_f32(tag, tol):
    # fp32 restatement vs the reference's fp32 eager path: agreement at the conditioning of the spline
    phi, x = U[f"rqs_{tag}_phi"], U[f"rqs_{tag}_x"]
    N, D = x.shape
    y, ladj = O.rqs_forward(x, phi.reshape(N, -1), 8, dtype=np.float32)
    np.testing.assert_allclose(y, U[f"rqs_{tag}_y32"], rtol=tol, atol=tol)
    np.testing.assert_allclose(ladj, U[f"rqs_{tag}_ladj32"], rtol=tol, atol=tol)


def test_rqs_edge_semantics():
    # strict '<' search, identity outside [-5, 5], ladj 0 there
    phi, x = U["rqs_s01_phi"], U["rqs_s01_x"]
    y, ladj = O.rqs_forward(x, phi.reshape(x.shape[0], -1), 8)
    for row, v in enumerate([-5.0, 5.0, np.float32(-5.0000005), None, 1e30, -1e30]):
        if v is None:
            continue
        assert y[row, 0] == np.float64(np.float32(v)) and ladj[row, 0] == 0.0, (row, v, y[row, 0], ladj[row, 0])


@pytest.mark.parametrize("K", [16, 5])
def test_rqs_shared_table(K):
    y, ladj = O.rqs_forward(U[f"rqs_shared{K}_x"], U[f"rqs_shared{K}_phi"], K)
    np.testing.assert_allclose(y, U[f"rqs_shared{K}_y64"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(ladj, U[f"rqs_shared{K}_ladj64"], rtol=1e-9, atol=1e-9)
    xi = O.rqs_inverse(U[f"rqs_shared{K}_x"], U[f"rqs_shared{K}_phi"], K)
    np.testing.assert_allclose(xi, U[f"rqs_shared{K}_xinv64"], rtol=1e-9, atol=1e-9)


def test_affine():
    phi, x = U["affine_phi"], U["affine_x"]
    y, ladj = O.affine_forward(x, phi.reshape(x.shape[0], -1))
    np.testing.assert_allclose(y, U["affine_y64"], **T64)
    np.testing.assert_allclose(ladj, U["affine_ladj64"], **T64)
    np.testing.assert_allclose(O.affine_inverse(x, phi.reshape(x.shape[0], -1)), U["affine_xinv64"], **T64)
    y32, ladj32 = O.affine_forward(x, phi.reshape(x.shape[0], -1), dtype=np.float32)
    np.testing.assert_allclose(y32, U["affine_y32"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(ladj32, U["affine_ladj32"], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("bound", [1, 11])
def test_softclip(bound):
    y, ladj = O.softclip_forward(U["softclip_x"], float(bound))
    np.testing.assert_allclose(y, U[f"softclip{bound}_y64"], **T64)
    np.testing.assert_allclose(ladj, U[f"softclip{bound}_ladj64"], **T64)
    np.testing.assert_allclose(O.softclip_inverse(U[f"softclip{bound}_y64"], float(bound)), U[f"softclip{bound}_xinv64"], rtol=1e-9, atol=1e-9)


def test_permutation_bit_exact():
    assert np.array_equal(O.permute(U["perm_x"], U["perm_order"]), U["perm_y"])
    assert np.array_equal(O.permute(U["perm_x"], U["perm_order"], inverse=True), U["perm_xinv"])


def test_rotation():
    from scipy.linalg import expm

    A = U["rot_A"].astype(np.float64)
    R = expm(A - A.T)
    np.testing.assert_allclose(R, U["rot_R64"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(O.rotate(U["perm_x"], R), U["rot_y64"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(O.rotate(U["perm_x"], R, transpose=True), U["rot_xinv64"], rtol=1e-10, atol=1e-10)


def test_diag_normal():
    lp = O.diag_normal_log_prob(U["dn_z"], U["dn_loc"], U["dn_scale"])
    np.testing.assert_allclose(lp, U["dn_lp64"], **T64)


@pytest.mark.parametrize("name", list(FLOW_CASES))
def test_flow_f64(name):
    g = load(f"flow_{name}")
    spec = O.flowspec_from_module(build_flow(name, g))
    c = g.get("c")
    z, ladj = spec.forward(g["x"], c)
    np.testing.assert_allclose(z, g["z64"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(ladj, g["ladj64"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(spec.log_prob(g["x"], c), g["log_prob64"], rtol=1e-9, atol=1e-8)
    if "zin" in g:
        n = g["zin"].shape[0]
        ci = None if c is None else (c if c.ndim == 1 else c[:n])
        np.testing.assert_allclose(spec.inverse(g["zin"], ci), g["xinv64"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("name", ["cfg1_maf", "cfg2_nsf", "nsf35_row", "nice35"])
def test_flow_f32(name):
    g = load(f"flow_{name}")
    spec = O.flowspec_from_module(build_flow(name, g))
    lp = spec.log_prob(g["x"], g.get("c"), dtype=np.float32)
    ref = g["log_prob32"].astype(np.float64)
    assert np.max(np.abs(lp - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-5


def test_inverse_and_log_prob_consistent():
    g = load("flow_nsf35_row")
    spec = O.flowspec_from_module(build_flow("nsf35_row", g))
    x, lp = spec.inverse_and_log_prob(g["zin"], g["c"])
    np.testing.assert_allclose(lp, spec.log_prob(x, g["c"]), rtol=1e-12, atol=1e-12)


# --------------------------------------------------------------------------- #
# circular spline flow pieces (NCSF): CircularShiftTransform, BoxUniform
# --------------------------------------------------------------------------- #

UN = load("units_ncsf")


@pytest.mark.parametrize("tag,bound", [("circ1", 1.0), ("circpi", float(np.pi))])
def test_circular_shift(tag, bound):
    x = UN["circ_x"]
    np.testing.assert_allclose(O.circular_shift(x, bound), UN[f"{tag}_y64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.circular_shift(x, bound), UN[f"{tag}_xinv64"], rtol=1e-12, atol=1e-12)  # its own inverse
    np.testing.assert_array_equal(O.circular_shift(x, bound, dtype=np.float32), UN[f"{tag}_y32"])  # fp32: bit-exact


def test_box_uniform_log_prob():
    lp = O.box_uniform_log_prob(UN["box_z"], UN["box_lower"], UN["box_upper"])
    ref = UN["box_lp64"]
    assert np.array_equal(np.isinf(lp), np.isinf(ref)) and np.isinf(ref).any() and np.isfinite(ref).any()
    np.testing.assert_allclose(lp[np.isfinite(ref)], ref[np.isfinite(ref)], rtol=1e-12)
