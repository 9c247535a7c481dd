"""Debug tool: copy-engine H2D bandwidth from pinned memory as a function of transfer size (decides the chunk size of the mapped
host transport) and whether two copy streams help."""
import torch
dev = torch.device("cuda", 0)
for mb in (0.25, 1.06, 2.1, 4.2, 8.5, 17, 34, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    reps = max(4, int(200e6 / n))
    with torch.cuda.stream(s):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        b.record(s)
    b.synchronize()
    t = a.elapsed_time(b) / reps
    print(f"H2D {mb:7.2f} MB: {t * 1e3:8.1f} us per copy, {n / t / 1e6:6.1f} GB/s")
# two streams, 8.5 MB each
n = int(8.5e6)
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for s in ss:
    s.wait_stream(torch.cuda.current_stream())
for r in range(20):
    for i in range(2):
        with torch.cuda.stream(ss[i]):
            ds[i].copy_(hs[i], non_blocking=True)
for s in ss:
    torch.cuda.current_stream().wait_stream(s)
b.record()
b.synchronize()
print(f"two streams x 8.5 MB: {40 * n / a.elapsed_time(b) / 1e6:6.1f} GB/s aggregate")
