"""Debug tool: per-CTA timeline of fused_wide_kernel (clock64 stamps), config 2.

    python scripts/wide_timeline.py [B] [bf16|bf16x3] [criteo|criteo_qr|twitter]
Slots (fused_wide.cuh): 28/30 kernel entry/exit; MMA warp 32+8it wait x_ready, 33+8it got it, 34+8it+l layer l issued, 39+8it tile
stored; epilogue warp 4: 64+16it+4l+2j accumulators of pair-tile j ready, +1 epilogue of j done; gather: 96+8it rows of round 0 in
registers, 97 X released, 98 x_ready arrive, 99 shallow ready.
"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
wl = sys.argv[3] if len(sys.argv) > 3 else "criteo"
bench.set_workload(wl.replace("_small", ""))
if wl.endswith("_small"):          # the same switches on the paper-Criteo cardinalities (L2-resident tables)
    from xsdeepfwfm_deprecated_b200.utils import workloads
    bench.SIZES = workloads.CRITEO_PAPER
m = bench.make_model(dev, prec, bench.SIZES)
plan = m._get_plan(); plan.ensure_image(m, prec)
Xi, Xv = bench.make_batches(dev, bench.SIZES, B, 4, seed=0)
out = torch.zeros(B, device=dev)
NCLK = 128
clk = torch.zeros(148 * NCLK, dtype=torch.int64, device=dev)
fn = lib.dfw_debug_set_fused_clock_buffer; fn.argtypes = [ctypes.c_void_p]; fn.restype = None
st = torch.cuda.current_stream().cuda_stream
def run(j=0):
    rc = lib.dfw_forward_fused(plan.model_ref, Xi[j].data_ptr(), bench.CATS, 1, Xv[j].data_ptr(), bench.NUM, 1, B, _lib.PRECISIONS[prec],
                               out.data_ptr(), None, None, st)
    _lib.check(rc, "dfw_forward_fused")
for j in range(3): run(j)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [run(j % 4) for j in range(8)]; e1.record(); torch.cuda.synchronize()
print(f"8 back-to-back launches: {e0.elapsed_time(e1) * 125:.1f} us per launch (CUDA events), {B / (e0.elapsed_time(e1) * 125e-6) / 1e6:.1f} M samples/s")
fn(clk.data_ptr()); run(3); torch.cuda.synchronize(); fn(None)
c = clk.cpu().numpy().reshape(148, NCLK).astype(np.float64)
used = c[:, 28] != 0
lead = used & (np.arange(148) % 2 == 0)
def med(slot, rows, ref=28):
    v = c[rows, slot] - c[rows, ref]
    v = v[c[rows, slot] != 0]
    return (np.median(v), v.min(), v.max()) if len(v) else (float("nan"),) * 3
print(f"B={B} {prec}: {used.sum()} CTAs; cycles since kernel entry, median / min / max")
L = 3
for it in range(4):
    if not (c[lead, 33 + 8 * it] != 0).any():
        break
    print(f"-- tile {it}")
    rows = [("gather: model state in shared memory (tile 0)", 102, used), ("gather: rows of the first sample in registers", 101 + 8 * it, used),
            ("gather: rows of both samples in registers", 96 + 8 * it, used), ("gather: chunks 0-3 of X released", 100 + 8 * it, used),
            ("gather: X released (x_free)", 97 + 8 * it, used),
            ("gather: x_ready arrive", 98 + 8 * it, used), ("MMA: waits for x_ready", 32 + 8 * it, lead), ("MMA: x_ready seen", 33 + 8 * it, lead)]
    for l in range(L):
        rows.append((f"MMA: layer {l + 1} issued", 34 + 8 * it + l, lead))
        if it < 2:
            for j in range(2):
                rows.append((f"epi: L{l + 1} pair-tile {j} accumulators ready", 64 + 16 * it + 4 * l + 2 * j, used))
                rows.append((f"epi: L{l + 1} pair-tile {j} done", 65 + 16 * it + 4 * l + 2 * j, used))
    rows += [("gather: shallow ready", 99 + 8 * it, used), ("tile stored", 39 + 8 * it, used)]
    for name, slot, r in rows:
        a, b_, c_ = med(slot, r)
        print(f"  {name:44s} {a:9.0f} {b_:9.0f} {c_:9.0f}")
a, b_, c_ = med(30, used)
print(f"kernel exit {a:9.0f} {b_:9.0f} {c_:9.0f}")
g0, g1 = c[used, 29], c[used, 31]
print(f"globaltimer: first CTA entry -> last CTA exit {g1.max() - g0.min():.0f} ns; CTA entry spread {g0.max() - g0.min():.0f} ns")
