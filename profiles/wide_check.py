"""Bring-up check of the wide fused layer kernel (hidden width 384 / 512): each case in its own
process (a protocol bug traps the context), fused vs per-layer path vs fp64 oracle; on failure the
kernel's watchdog report is decoded.  Usage: python profiles/wide_check.py [case ...]"""
import subprocess
import sys

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")

CASES = {
    "maf32_h512x4": "zuko.flows.MAF(32, 0, transforms=2, hidden_features=[512] * 4)",
    "nsf64c16_k16_h512": "zuko.flows.NSF(64, 16, transforms=2, bins=16, hidden_features=[512] * 3)",
    "nsf24_k8_h512": "zuko.flows.NSF(24, 0, transforms=2, bins=8, hidden_features=[512, 512])",
    "nsf10c3_h384": "zuko.flows.NSF(10, 3, transforms=3, bins=8, hidden_features=[384, 384])",
    "maf100c28_h512": "zuko.flows.MAF(100, 28, transforms=2, hidden_features=[512] * 2)",
}
CODES = {0x10: "W producer: w_empty", 0x20: "issuer: s_ready", 0x30: "scout: d_empty", 0x31: "scout: a_ready (unread)",
         0x32: "scout: a_ready", 0x33: "scout: w_full", 0x40: "epi hidden: d_full", 0x41: "epi hidden: a_free",
         0x42: "epi out: d_full", 0x43: "epi tile end: a_free"}


def dump_watchdog():
    import numpy as np
    from zuko_b200 import _engine as E

    buf = np.zeros(1024, np.uint32)
    n = E.lib().zk_debug_watchdog_read(buf.ctypes.data, 1024)
    print(f"  watchdog words={n} reports={buf[0]}")
    for cta in range(2):
        for w in range(20):
            s = buf[8 + (cta * 32 + w) * 8 :][:8]
            if s[0]:
                print(f"   cta{cta} warp{w:2d}: {CODES.get(int(s[0]) & 0xFFFF, hex(int(s[0])))}  a={int(s[1])} b={int(s[2])} tid={int(s[3])}")


def run_case(name, B):
    import numpy as np
    import torch

    import zuko_b200 as zuko  # noqa: F401
    from oracle import oracle as O
    from zuko_b200 import _engine as E

    torch.manual_seed(11)
    flow_cpu = eval(CASES[name]).eval()
    spec = O.flowspec_from_module(flow_cpu)
    D = flow_cpu.base.loc.shape[0]
    C = flow_cpu.transform.transforms[0].context
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, D, generator=g)
    c = torch.randn(B, C, generator=g) if C else None
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    flow = eval(CASES[name])
    flow.load_state_dict(flow_cpu.state_dict())
    flow = flow.to(dev)
    xd, cd = x.to(dev), (None if c is None else c.to(dev))
    try:
        flow(None if cd is None else cd[:1]).log_prob(xd[:1])  # packs the handles (pack kernels count as launches)
        n0 = E.lib().zk_launch_count()
        lp_f = flow(cd).log_prob(xd)
        torch.cuda.synchronize()
        launches = E.lib().zk_launch_count() - n0
        z_f, ladj_f = flow(cd).transform.call_and_ladj(xd)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"{name} B={B}: FAILED {type(e).__name__}: {str(e)[:200]}")
        dump_watchdog()
        return 1
    prev = E.lib().zk_set_fused_layers(0)
    lp_u = flow(cd).log_prob(xd)
    z_u, ladj_u = flow(cd).transform.call_and_ladj(xd)
    E.lib().zk_set_fused_layers(prev)
    nb = min(B, 2048)
    ref = spec.log_prob(x[:nb].numpy(), None if c is None else c[:nb].numpy())
    rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))  # noqa: E731
    e_f = rel(lp_f[:nb].cpu().numpy().astype(np.float64), ref)
    e_u = rel(lp_u[:nb].cpu().numpy().astype(np.float64), ref)
    d_fu = float((lp_f - lp_u).abs().max())
    d_z = float((z_f - z_u).abs().max())
    ok = e_f < 1e-5 and d_z < 1e-4 and launches == len(flow.transform.transforms)
    print(f"{name} B={B}: launches={launches} fused-vs-oracle {e_f:.2e}  unfused-vs-oracle {e_u:.2e}  |lp_f-lp_u| {d_fu:.2e} |z_f-z_u| {d_z:.2e} {'OK' if ok else 'MISMATCH'}")
    return 0 if ok else 2


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        rc = 0
        for B in (1, 300, 256 * 148 + 77, 100000):
            rc = max(rc, run_case(sys.argv[2], B))
            if rc == 1:
                break  # the kernel trapped: the context is gone
        sys.exit(rc)
    names = [a for a in sys.argv[1:] if a in CASES] or list(CASES)
    bad = 0
    for name in names:
        r = subprocess.run([sys.executable, __file__, "--one", name], timeout=600)
        bad += r.returncode != 0
    print("wide_check:", "ALL OK" if bad == 0 else f"{bad} failures")
    sys.exit(1 if bad else 0)
