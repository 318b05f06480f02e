"""Host <-> device copy bandwidth of the box (pinned memory), one and two streams: the ceiling of the e2e number."""
import torch, time
dev = torch.device('cuda:0')
n = 100 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
for label, fn in (("H2D 100 MB, one stream", lambda: d.copy_(h, non_blocking=True)), ("D2H 100 MB, one stream", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label}: {ms:.2f} ms -> {n / ms / 1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"H2D 100 MB split over two streams: {ms:.2f} ms -> {n / ms / 1e6:.1f} GB/s")
