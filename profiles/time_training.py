"""Times training-type steps of the BASELINE configs with CUDA events (one JSON line each):

    python profiles/time_training.py            # cfg2 / cfg3 / cfg5 maximum-likelihood steps + a cfg2 reverse-KL step

* maximum likelihood: ``(-flow(c).log_prob(x).mean()).backward()``  (zk_flow_log_prob + zk_flow_backward)
* reverse KL:         ``x, lq = flow(c).rsample_and_log_prob(...); (lq - log p*(x)).mean().backward()``
                      (zk_flow_inverse + zk_flow_inverse_backward)
"""

import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import zuko_b200 as zuko  # noqa: E402
from zuko_b200 import _engine as E  # noqa: E402

dev = torch.device("cuda:0")
CONFIGS = {
    "cfg2": (lambda: zuko.flows.NSF(16, 8, transforms=4, bins=8, hidden_features=[256] * 3), 16, 8, 1 << 18),
    "cfg3": (lambda: zuko.flows.MAF(32, 0, transforms=8, hidden_features=[512] * 4), 32, 0, 1 << 17),
    "cfg5": (lambda: zuko.flows.NSF(64, 16, transforms=8, bins=16, hidden_features=[512] * 3), 64, 16, 1 << 16),
}


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    n0 = E.lib().zk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (E.lib().zk_launch_count() - n0) / iters


for name, (build, D, C, rows) in CONFIGS.items():
    torch.manual_seed(0)
    flow = build().to(dev)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, D, generator=g).to(dev)
    c = torch.randn(rows, C, generator=g).to(dev) if C else None

    def ml_step():
        for p in flow.parameters():
            p.grad = None
        (-flow(c).log_prob(x).mean()).backward()

    ms, launches = timed(ml_step, 3)
    with torch.no_grad():
        fwd, _ = timed(lambda: flow(c).log_prob(x), 3)
    lins = flow.transform.transforms[0].hyper._linears()
    flops = 2.0 * sum(m.weight.numel() for m in lins) * rows * len(flow.transform.transforms)
    print(json.dumps({"step": "maximum likelihood (log_prob forward + backward)", "config": name, "rows": rows, "ms_per_step": ms,
                      "ms_forward_only": fwd, "samples_per_s": rows / (ms * 1e-3), "launches_per_step": launches,
                      "backward_tflops_algorithmic": 3 * flops / max(ms - fwd, 1e-9) / 1e9}))
    if name == "cfg2":
        r2 = 1 << 15
        cc = None if c is None else c[:r2]

        def kl_step():
            for p in flow.parameters():
                p.grad = None
            xs, lq = flow(cc).rsample_and_log_prob(() if cc is not None else (r2,))
            target = -0.5 * (xs**2).sum(-1)
            (lq - target).mean().backward()

        ms, launches = timed(kl_step, 2)
        with torch.no_grad():
            smp, _ = timed(lambda: flow(cc).rsample_and_log_prob(() if cc is not None else (r2,)), 2)
        print(json.dumps({"step": "reverse KL (rsample_and_log_prob + backward through the sampler)", "config": name, "rows": r2,
                          "ms_per_step": ms, "ms_sampling_only": smp, "samples_per_s": r2 / (ms * 1e-3), "launches_per_step": launches}))
