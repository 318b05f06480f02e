"""Golden vectors (values + gradients) of flows whose conditioners use a non-ReLU activation
(zuko/nn.py:160-192, 258-318: `activation=`), from the UNMODIFIED reference.  Separate from
make_golden.py so that the existing files are not rewritten.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_act.py
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
from zuko.flows import MAF, NICE, NSF  # noqa: E402

CASES = {
    "act_maf_elu": (lambda: MAF(5, 2, transforms=2, hidden_features=[32, 32], activation=nn.ELU), ("batch", 2)),
    "act_nsf_tanh": (lambda: NSF(4, 0, transforms=2, hidden_features=[64, 64], activation=nn.Tanh), None),
    "act_nsf_silu": (lambda: NSF(3, 2, transforms=2, hidden_features=[32], activation=nn.SiLU), ("row", 2)),
    "act_maf_gelu": (lambda: MAF(4, 0, transforms=2, hidden_features=[64, 64], activation=nn.GELU), None),
    "act_nice_lrelu": (lambda: NICE(4, 2, hidden_features=[32], activation=nn.LeakyReLU), ("batch", 2)),
    "act_maf_softplus": (lambda: MAF(3, 0, transforms=2, hidden_features=[16], activation=nn.Softplus), None),
    "act_maf_sigmoid": (lambda: MAF(3, 1, transforms=2, hidden_features=[16, 16], activation=nn.Sigmoid), ("batch", 1)),
}


def main():
    for name, (build, ctx) in CASES.items():
        MG.flow_case(name, build, 128, ctx, store=True, inverse_rows=32)
        MGG.grad_case(name, build, 128)


if __name__ == "__main__":
    main()
