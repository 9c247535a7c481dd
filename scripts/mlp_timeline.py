"""Debug tool: per-CTA timeline of mlp_tc_kernel (clock64 stamps), B=4096 config 2, bf16."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = bench.make_model(dev, "bf16", synth.CRITEO_PAPER)
plan = m._get_plan(); plan.ensure_image(m, "bf16")
Eb = (torch.randn(B + 128, 392, device=dev) * 0.1).bfloat16(); sh = torch.zeros(B + 128, device=dev)
out = torch.zeros(B, device=dev); ws = torch.zeros(8192, dtype=torch.uint8, device=dev)
nc = min((B + 127) // 128, 148)
clk = torch.zeros(nc * 32, dtype=torch.int64, device=dev)
fn = lib.dfw_debug_set_mlp_clock_buffer; fn.argtypes = [ctypes.c_void_p]; fn.restype = None
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.dfw_mlp_bf16(plan.model_ref, Eb.data_ptr(), 392, B, sh.data_ptr(), ws.data_ptr(), ws.numel(), out.data_ptr(), None, st)
    assert rc == 0
for _ in range(3): run()
torch.cuda.synchronize()
fn(clk.data_ptr()); run(); torch.cuda.synchronize(); fn(None)
c = clk.cpu().numpy().reshape(nc, 32).astype(np.float64)
t0 = c[:, 0:1]
rel = c - t0
names = {1: "X landed"}
for l in range(3):
    for q in range(2):
        names[2 + 2 * l + q] = f"L{l+1} p{q} mma issued"
        names[10 + 4 * l + 2 * q] = f"L{l+1} p{q} acc ready"
        names[11 + 4 * l + 2 * q] = f"L{l+1} p{q} epi done"
order = [1] + [k for l in range(3) for k in (2 + 2 * l, 10 + 4 * l, 3 + 2 * l, 11 + 4 * l, 12 + 4 * l, 13 + 4 * l)]
print("cycles since MMA-thread start (median over CTAs / max)")
for k in order:
    print(f"  {names[k]:16s} {np.median(rel[:, k]):9.0f} {rel[:, k].max():9.0f}")
