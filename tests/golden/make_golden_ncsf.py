"""Golden vectors of the circular spline flow (NCSF, zuko/flows/spline.py:65-117) and of
CircularShiftTransform / BoxUniform, from the UNMODIFIED reference.  Separate from make_golden.py so
that the existing files are not rewritten.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ncsf.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parent))

import make_golden as MG  # noqa: E402  (flow_case; imports the reference)
import make_golden_grad as MGG  # noqa: E402
from zuko.distributions import BoxUniform  # noqa: E402
from zuko.flows import NCSF  # noqa: E402
from zuko.transforms import CircularShiftTransform  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    # x ~ N(0, 1.5^2): mostly inside [-pi, pi[, with a tail that wraps around
    MG.flow_case("ncsf34", lambda: NCSF(3, 4, hidden_features=[32, 32]), 256, ("batch", 4), store=True, x_scale=1.5, inverse_rows=64)
    MGG.grad_case("ncsf34", lambda: NCSF(3, 4, hidden_features=[32, 32]), 128)
    out = {}
    x = MG.gen(900, 64, 5, scale=4.0)
    x[0, :4] = torch.tensor([-np.pi, np.pi, 0.0, 2 * np.pi], dtype=torch.float32)
    for b in (1.0, float(np.pi)):
        t = CircularShiftTransform(bound=b)
        tag = "circ1" if b == 1.0 else "circpi"
        out[f"{tag}_y32"] = t(x).numpy()
        out[f"{tag}_y64"] = t(x.double()).numpy()
        out[f"{tag}_xinv64"] = t.inv(x.double()).numpy()
    out["circ_x"] = x.numpy()
    lower, upper = torch.tensor([-1.0, 0.0, -2.0]), torch.tensor([1.0, 3.0, 2.5])
    z = MG.gen(901, 64, 3, scale=1.5)
    z[0] = torch.tensor([-1.0, 0.0, -2.0])  # lower bounds are inclusive
    z[1] = torch.tensor([1.0, 1.0, 1.0])  # upper bounds are exclusive
    out["box_z"], out["box_lower"], out["box_upper"] = z.numpy(), lower.numpy(), upper.numpy()
    out["box_lp64"] = BoxUniform(lower.double(), upper.double()).log_prob(z.double()).numpy()
    np.savez_compressed(OUT / "units_ncsf.npz", **out)
    print("units_ncsf: done")


if __name__ == "__main__":
    main()
