"""Debug tool: per-CTA timeline of fused_forward_kernel (clock64 stamps), config 2, B=4096 by default.

    python scripts/fused_timeline.py [B] [bf16|bf16x3] [criteo|criteo_qr|twitter]
"""
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
bench.set_workload(sys.argv[3] if len(sys.argv) > 3 else "criteo")
m = bench.make_model(dev, prec, bench.SIZES)
plan = m._get_plan(); plan.ensure_image(m, prec)
Xi, Xv = bench.make_batches(dev, bench.SIZES, B, 4, seed=0)
out = torch.zeros(B, device=dev)
nc = min((B + 31) // 32, 148)
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
fn(clk.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(3); run(0); run(1); run(2); e1.record(); torch.cuda.synchronize(); fn(None)
print(f"4 back-to-back launches: {e0.elapsed_time(e1) * 250:.1f} us per launch (CUDA events)")
c = clk.cpu().numpy().reshape(148, NCLK)[:nc].astype(np.float64)
if (c[:, 0] == 0).any():          # pair kernel: only the leader CTA of a pair has MMA threads; use the leaders' rows
    c = c[c[:, 0] != 0]
    nc = len(c)
t0 = c[:, 0:1]          # MMA thread reaches the x_ready wait
rel = c - t0
names = {20: "gather done (E block)", 21: "X written (bf16)", 1: "x_ready seen by MMA", 22: "shallow done", 16: "tile done"}
for l in range(3):
    names[2 + l] = f"L{l+1} all MMAs issued"
    names[8 + 2 * l] = f"L{l+1} first acc ready"
    names[9 + 2 * l] = f"L{l+1} epilogue done"
order = [20, 21, 1, 8, 2, 9, 10, 3, 11, 12, 4, 13, 22, 16]
print(f"B={B} {prec}: cycles relative to the MMA thread's start (median over {nc} CTAs / min / max)")
for k in order:
    print(f"  {names[k]:24s} {np.median(rel[:, k]):9.0f} {rel[:, k].min():9.0f} {rel[:, k].max():9.0f}")

print("gather group phases (median): start, image+idx regs, idx in smem, rows issued, E complete, interact start, phase D done")
print("  " + " ".join(f"{np.median(rel[:, 96 + k]):9.0f}" for k in range(7)))
if c[:, 96 + 8].any():
    print("operand write (median): values in registers, group sync, joined set-up, columns stored, proxy fence done")
    print("  " + " ".join(f"{np.median(rel[:, 96 + k]):9.0f}" for k in range(8, 13)))

print(f"kernel entry -> MMA thread start: median {np.median(-rel[:, 28]):.0f} cycles; MMA start -> exit: {np.median(rel[:, 30]):.0f} cycles")
g0, g1 = c[:, 29], c[:, 31]
print(f"globaltimer: first CTA entry -> last CTA exit {g1.max() - g0.min():.0f} ns; CTA entry spread {g0.max() - g0.min():.0f} ns; per-CTA lifetime median {np.median(g1 - g0):.0f} ns")

if c[:, 40].any():
    print("layer-1 epilogue of warp 4 (median): per pair-tile: start, TMEM loaded, stores issued, proxy fence done")
    for j in range(2):
        print("  tile", j, " ".join(f"{np.median(rel[:, 40 + 4 * j + k]):9.0f}" for k in range(4)))
