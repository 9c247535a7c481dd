"""Debug tool: what do the per-step host<->device copies of the e2e path cost by themselves?"""
import time, torch
dev = torch.device("cuda", 0)
B, n = 4096, 32
hXi = torch.randint(0, 4, (n, B, 26), dtype=torch.int64).pin_memory(); hXv = torch.rand(n, B, 13).pin_memory(); hout = torch.empty(n, B).pin_memory()
hAll = torch.empty(n, B * 26 * 8 + B * 13 * 4, dtype=torch.uint8).pin_memory()
dXi = [torch.empty(B, 26, dtype=torch.int64, device=dev) for _ in range(3)]; dXv = [torch.empty(B, 13, device=dev) for _ in range(3)]
dAll = [torch.empty(B * 26 * 8 + B * 13 * 4, dtype=torch.uint8, device=dev) for _ in range(3)]
dout = [torch.zeros(B, device=dev) for _ in range(3)]
streams = [torch.cuda.Stream(dev) for _ in range(3)]
def run(mode, ns):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for rep in range(10):
        for i in range(n):
            k = i % ns
            with torch.cuda.stream(streams[k]):
                if mode == "split":
                    dXi[k].copy_(hXi[i], non_blocking=True); dXv[k].copy_(hXv[i], non_blocking=True)
                elif mode == "merged":
                    dAll[k].copy_(hAll[i], non_blocking=True)
                elif mode == "xi_only":
                    dXi[k].copy_(hXi[i], non_blocking=True)
                if mode != "xi_only_nod2h":
                    hout[i].copy_(dout[k], non_blocking=True)
        torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / (10 * n)
    print(f"{mode:8s} streams={ns}: {t * 1e6:6.1f} us/step  ({(B * 26 * 8 + B * 13 * 4) / t / 1e9:5.1f} GB/s H2D)")
for mode in ("split", "merged", "xi_only"):
    for ns in (1, 3):
        run(mode, ns)
# large-copy PCIe bandwidth for reference
big_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    big_d.copy_(big_h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    big_d.copy_(big_h, non_blocking=True)
torch.cuda.synchronize()
print(f"256 MiB H2D copies: {5 * (256 << 20) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
