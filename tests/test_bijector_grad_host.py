"""CPU tests of the engine's per-pair backward math (zuko_b200/csrc/bijector_grad.cuh): the header
is host/device code; here g++ compiles it into a small harness and the result is held against
gradients torch.autograd produced on the unmodified reference (tests/golden/grad_units.npz).
The GPU kernel (uni_bwd_kernel) executes exactly these functions."""

import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from cases import load

ROOT = Path(__file__).resolve().parent.parent
GU = load("grad_units")
U = load("units")
_P = ctypes.c_void_p


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("native") / "libbijgrad_host.so"
    src = ROOT / "tests" / "native" / "bijector_grad_host.cpp"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", str(src), "-o", str(so), "-lm"], check=True)
    return ctypes.CDLL(str(so))


def _ptr(a):
    return a.ctypes.data_as(_P)


def _close(ours, ref, tol, what):
    # relative to the scale of each row's gradient: the spline's parameter gradients span decades
    scale = np.maximum(np.abs(ref).max(axis=-1, keepdims=True) if ref.ndim > 1 else np.abs(ref), 1.0)
    err = np.abs(ours - ref) / scale
    assert err.max() <= tol, f"{what}: max err {err.max():.3e} (bar {tol:.1e}) at {np.unravel_index(err.argmax(), err.shape)}"


@pytest.mark.parametrize("tag,tol", [("s01", 2e-5), ("s1", 5e-5), ("s3", 2e-3)])
def test_rqs_pair_backward(harness, tag, tol):
    x = GU[f"rqs_{tag}_x"].astype(np.float32)
    phi = np.ascontiguousarray(U[f"rqs_{tag}_phi"].astype(np.float32))
    gy, gl = GU[f"rqs_{tag}_gy"].astype(np.float32), GU[f"rqs_{tag}_gl"].astype(np.float32)
    n = x.size
    gx = np.empty(n, np.float32)
    gphi = np.empty((n, 23), np.float32)
    harness.rqs_backward_pairs(_ptr(x), _ptr(phi), _ptr(gy), _ptr(gl), ctypes.c_int64(n), 8, ctypes.c_float(5.0),
                               ctypes.c_float(1e-3), _ptr(gx), _ptr(gphi))
    ok = (np.abs(x) != 5.0).reshape(-1)  # on-knot element: see tests/test_oracle_grad.py
    ref_gx, ref_gphi = GU[f"rqs_{tag}_gx"].reshape(-1), GU[f"rqs_{tag}_gphi"].reshape(n, 23)
    _close(gx[ok], ref_gx[ok], tol, "gx")
    _close(gphi[ok], ref_gphi[ok], tol, "gphi")
    # compile-time K and in-place output give the same numbers
    phi2 = phi.copy().reshape(n, 23)
    gx2 = np.empty(n, np.float32)
    harness.rqs_backward_pairs_k8_inplace(_ptr(x), _ptr(phi2), _ptr(gy), _ptr(gl), ctypes.c_int64(n), ctypes.c_float(5.0),
                                          ctypes.c_float(1e-3), _ptr(gx2))
    np.testing.assert_array_equal(gx2, gx)
    np.testing.assert_array_equal(phi2, gphi)


@pytest.mark.parametrize("K", [16, 5])
def test_rqs_pair_backward_other_bins(harness, K):
    x = U[f"rqs_shared{K}_x"].astype(np.float32)
    N, D = x.shape
    P = 3 * K - 1
    phi = np.ascontiguousarray(np.broadcast_to(U[f"rqs_shared{K}_phi"].astype(np.float32), (N, D, P)))
    gy, gl = GU[f"rqs_shared{K}_gy"].astype(np.float32), GU[f"rqs_shared{K}_gl"].astype(np.float32)
    gx = np.empty(N * D, np.float32)
    gphi = np.empty((N * D, P), np.float32)
    harness.rqs_backward_pairs(_ptr(x), _ptr(phi), _ptr(gy), _ptr(gl), ctypes.c_int64(N * D), K, ctypes.c_float(5.0),
                               ctypes.c_float(1e-3), _ptr(gx), _ptr(gphi))
    _close(gx, GU[f"rqs_shared{K}_gx"].reshape(-1), 5e-5, "gx")
    # shared table: the reference's gradient is the sum over the batch
    total = gphi.reshape(N, D, P).astype(np.float64).sum(0)
    _close(total, GU[f"rqs_shared{K}_gphi"], 5e-5, "gphi (summed over the batch)")


def test_affine_pair_backward(harness):
    x = U["affine_x"].astype(np.float32)
    phi = np.ascontiguousarray(U["affine_phi"].astype(np.float32))
    gy, gl = GU["affine_gy"].astype(np.float32), GU["affine_gl"].astype(np.float32)
    n = x.size
    gx = np.empty(n, np.float32)
    gphi = np.empty((n, 2), np.float32)
    harness.affine_backward_pairs(_ptr(x), _ptr(phi), _ptr(gy), _ptr(gl), ctypes.c_int64(n), ctypes.c_float(1e-3), _ptr(gx), _ptr(gphi))
    _close(gx, GU["affine_gx"].reshape(-1), 1e-5, "gx")
    _close(gphi, GU["affine_gphi"].reshape(n, 2), 1e-5, "gphi")
