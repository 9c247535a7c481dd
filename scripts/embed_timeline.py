"""Debug tool: per-CTA phase timeline of embed_fwfm_kernel (clock64 stamps), B=4096 config 2."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = bench.make_model(dev, "fp32", synth.CRITEO_PAPER)
Xi, Xv = bench.make_batches(dev, synth.CRITEO_PAPER, B, 8, 0)
plan = m._get_plan()
E = torch.zeros(B + 128, 392, device=dev); sh = torch.zeros(B + 128, device=dev)
nc = (B + 15) // 16
clk = torch.zeros(nc * 8, dtype=torch.int64, device=dev)
fn = lib.dfw_debug_set_clock_buffer; fn.argtypes = [ctypes.c_void_p]; fn.restype = None
st = torch.cuda.current_stream().cuda_stream
def run(j):
    rc = lib.dfw_embed_fwfm(plan.model_ref, Xi[j].data_ptr(), 26, 1, Xv[j].data_ptr(), 13, 1, B, E.data_ptr(), 392, None, 0, sh.data_ptr(), None, st)
    assert rc == 0
for j in range(4): run(j)
torch.cuda.synchronize()
fn(clk.data_ptr()); run(5); torch.cuda.synchronize(); fn(None)
c = clk.cpu().numpy().reshape(nc, 8)
d = np.diff(c[:, :7], axis=1)
names = ["A:image+inputs", "A2:idx check", "B:issue rows", "B2:wait+fixup", "C:stream E", "D:fwfm"]
print("per-CTA phase cycles  median / p90 / max")
for i, n in enumerate(names):
    print(f"  {n:16s} {np.median(d[:, i]):8.0f} {np.percentile(d[:, i], 90):8.0f} {d[:, i].max():8.0f}")
tot = c[:, 6] - c[:, 0]
print(f"  total            {np.median(tot):8.0f} {np.percentile(tot, 90):8.0f} {tot.max():8.0f}")
g = c[:, 7]
print("CTA start spread (ns):", int(g.max() - g.min()), " first-to-last-start; kernel span >= ", int(g.max() - g.min()) + int(tot.max() / 1.9))
