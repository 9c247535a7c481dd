"""Debug tool: where does the time between consecutive fused launches go?  Captures 6 launches (each with its own clock buffer)
into one CUDA graph and prints, per launch, entry/exit in the global ns timer."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
m = bench.make_model(dev, prec, synth.CRITEO_PAPER)
plan = m._get_plan(); plan.ensure_image(m, prec)
NL = 6
Xi, Xv = bench.make_batches(dev, synth.CRITEO_PAPER, B, NL, seed=0)
out = torch.zeros(NL, B, device=dev)
NCLK = 128
clk = torch.zeros(NL, 148 * NCLK, dtype=torch.int64, device=dev)
fn = lib.dfw_debug_set_fused_clock_buffer; fn.argtypes = [ctypes.c_void_p]; fn.restype = None
stream = torch.cuda.Stream(dev)
def run(j):
    rc = lib.dfw_forward_fused(plan.model_ref, Xi[j].data_ptr(), 26, 1, Xv[j].data_ptr(), 13, 1, B, _lib.PRECISIONS[prec],
                               out[j].data_ptr(), None, None, stream.cuda_stream)
    _lib.check(rc, "dfw_forward_fused")
with torch.cuda.stream(stream):
    for j in range(3): run(j)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=stream):
    for j in range(NL):
        fn(clk[j].data_ptr()); run(j)
fn(None)
with torch.cuda.stream(stream):
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); g.replay(); e1.record(stream); torch.cuda.synchronize()
print(f"graph of {NL} launches: {e0.elapsed_time(e1) * 1e3 / NL:.1f} us per launch")
nc = min((B + 31) // 32, 148)
c = clk.cpu().numpy().reshape(NL, 148, NCLK)[:, :nc]
t0 = c[0, :, 29].min()
for j in range(NL):
    ent, ex = c[j, :, 29], c[j, :, 31]
    gap = (ent.min() - c[j - 1, :, 31].max()) if j else 0
    print(f"launch {j}: first entry {ent.min() - t0:7d} ns  last entry {ent.max() - t0:7d}  first exit {ex.min() - t0:7d}  last exit {ex.max() - t0:7d}  "
          f"lifetime {np.median(ex - ent):6.0f}  gap after previous {gap:6d} ns")
