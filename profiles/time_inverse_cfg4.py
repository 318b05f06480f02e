"""cfg4 (NSF(64, K16, T8) rsample 2^20): the dimension-sequential inverse kernel under several CTA
geometries (ZK_INV_GEOM=TxR: threads x samples per thread), each in its own process."""
import os, subprocess, sys
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1 and sys.argv[1] == '--one':
    import torch
    import zuko_b200 as zuko
    torch.manual_seed(0)
    dev = torch.device('cuda:0')
    flow = zuko.flows.NSF(64, 0, transforms=8, bins=16).to(dev)  # BASELINE config 4
    N = 1 << 20
    z = torch.randn(N, 64, device=dev)
    with torch.no_grad():
        t = flow().transform
        t.inv(z[:1024]); torch.cuda.synchronize()
        x = t.inv(z); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): x = t.inv(z)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        back = (t(x[:65536]) - z[:65536]).abs().max().item()
    print(f"cfg4 inverse N={N} geom={os.environ.get('ZK_INV_GEOM', 'default')}: {ms:.1f} ms -> {N / ms * 1e3:.3e} samples/s, round trip {back:.1e}")
else:
    for geom in (sys.argv[1:] or ['default', '128x1', '256x1', '128x2', '64x2', '96x2']):
        env = dict(os.environ)
        if geom != 'default': env['ZK_INV_GEOM'] = geom
        subprocess.run([sys.executable, __file__, '--one'], env=env, timeout=300)
