"""Golden vectors (values, inverse, gradients of both directions) of flows whose conditioners use
residual blocks (MaskedMLP(residual=True), zuko/nn.py:195-199, 297-309; exercised by the reference's
tests/test_nn.py:39-60), from the UNMODIFIED reference.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_res.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch.nn as nn

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parent))

import make_golden as MG  # noqa: E402
import make_golden_grad as MGG  # noqa: E402
import make_golden_inv_grad as MGI  # noqa: E402
from zuko.flows import MAF, NSF  # noqa: E402

CASES = {
    "res_nsf_relu": (lambda: NSF(5, 2, transforms=2, residual=True, hidden_features=[32, 32]), ("batch", 2)),
    "res_maf_elu": (lambda: MAF(4, 0, transforms=2, residual=True, hidden_features=[24], activation=nn.ELU), None),
    "res_nsf_mixed": (lambda: NSF(3, 1, transforms=2, residual=True, hidden_features=[16, 32, 32]), ("row", 1)),
}


def main():
    for name, (build, ctx) in CASES.items():
        MG.flow_case(name, build, 128, ctx, store=True, inverse_rows=32)
        MGG.grad_case(name, build, 128)
        MGI.inv_grad_case(name, build, 32)


if __name__ == "__main__":
    main()
