"""GRADIENT golden vectors of the INVERSE direction (reparameterised sampling) from
the UNMODIFIED reference:  for a fixed z (the `zin` of flow_*.npz)

* ``L1 = <w, x>``                      with ``x = flow(c).transform.inv(z)``                     (rsample)
* ``L2 = <w, x> + <wl, lp>``           with ``x, ladj = flow(c).transform.inv.call_and_ladj(z)``,
                                       ``lp = flow(c).base.log_prob(z) - ladj``  (rsample_and_log_prob,
                                       zuko/distributions.py:129-138)

differentiated by torch.autograd in fp64 (through the reference's `passes` inverse sweeps,
zuko/transforms.py:994-1000) w.r.t. z, c and every parameter.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_inv_grad.py
"""

from __future__ import annotations

import copy
import sys
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parent))

import make_golden_act as MA  # noqa: E402
import make_golden_grad as MGG  # noqa: E402
from zuko.flows import MAF, NCSF, NICE, NSF  # noqa: E402

OUT = Path(__file__).resolve().parent


def inv_grad_case(name: str, build, rows: int, *, seed=0, full=True):
    src = np.load(OUT / f"flow_{name}.npz")
    if seed is not None:
        torch.manual_seed(seed)
    flow = build().eval()
    f64 = copy.deepcopy(flow).double()
    rows = min(rows, src["zin"].shape[0])
    z = torch.from_numpy(src["zin"][:rows]).double()
    if name.startswith("ncsf"):  # a circular flow is a bijection of [-pi, pi[ only: keep z inside
        z = z * (3.0 / max(3.0, float(z.abs().max()) + 1e-3))
    c = None
    if "c" in src.files:
        c = torch.from_numpy(src["c"])
        c = (c if c.dim() == 1 else c[:rows]).double()
    B, D = z.shape
    with torch.no_grad():
        x0 = f64(c).transform.inv(z)
    if not name.startswith("ncsf"):
        assert np.allclose(x0.numpy(), src["xinv64"][:rows], rtol=1e-10, atol=1e-10), name
    w = MGG.gen(9501, B, D).double()
    wl = MGG.gen(9502, B).double()
    info = {"rows": np.int64(rows), "w": w.numpy(), "wl": wl.numpy(), "x64": x0.numpy(), "z": z.numpy()}

    def l1(f, zz, cc):
        return (w * f(cc).transform.inv(zz)).sum()

    def l2(f, zz, cc):
        d = f(cc)
        x, ladj = d.transform.inv.call_and_ladj(zz)
        lp = d.base.log_prob(zz) - ladj
        return (w * x).sum() + (wl * lp).sum()

    info.update(MGG.grads_of(f64, z, c, l1, full, "inv/"))
    info.update(MGG.grads_of(f64, z, c, l2, full, "invlp/"))
    with torch.no_grad():
        d = f64(c)
        x, ladj = d.transform.inv.call_and_ladj(z)
        info["lp64"] = (d.base.log_prob(z) - ladj).numpy()
    np.savez_compressed(OUT / f"invgrad_{name}.npz", **info)
    print(f"invgrad_{name}: rows={rows} D={D} |gz|max={np.abs(info['inv/gx']).max():.3e}")


def main():
    torch.set_num_threads(8)
    inv_grad_case("cfg1_maf", lambda: MAF(4, 0, transforms=2, hidden_features=[32, 32]), 64)
    inv_grad_case("cfg2_nsf", lambda: NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 16, full=False)
    inv_grad_case("nsf35_row", lambda: NSF(3, 5), 64)
    inv_grad_case("maf35_batch", lambda: MAF(3, 5), 64)
    inv_grad_case("nice35", lambda: NICE(3, 5), 64)
    inv_grad_case("nsf5_passes2", lambda: NSF(5, 0, passes=2, hidden_features=[32, 32]), 32)
    inv_grad_case("maf5_randperm", lambda: MAF(5, 2, randperm=True, hidden_features=[24]), 32, seed=7)
    inv_grad_case("nsf1_elementwise", lambda: NSF(1, 3, hidden_features=[16]), 32)
    inv_grad_case("ncsf34", lambda: NCSF(3, 4, hidden_features=[32, 32]), 64)
    inv_grad_case("act_maf_elu", MA.CASES["act_maf_elu"][0], 32)
    inv_grad_case("act_nsf_tanh", MA.CASES["act_nsf_tanh"][0], 32)
    # user-composed stacks (softclip / permutation / rotation / coupling / element-wise members)
    import make_golden_grad

    _orig = make_golden_grad.grad_case
    try:
        make_golden_grad.grad_case = lambda name, build, rows, seed=0, full=True, w_scale=1.0: inv_grad_case(name, build, rows, seed=seed, full=full)
        make_golden_grad.composed_grad_cases()
    finally:
        make_golden_grad.grad_case = _orig


if __name__ == "__main__":
    main()
