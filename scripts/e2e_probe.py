"""Debug tool: time the streamed host API (dfw_forward_host_stream) alone."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B, nh = 4096, 32
m = bench.make_model(dev, prec, synth.CRITEO_PAPER)
plan = m._get_plan(); plan.ensure_image(m, prec)
P = _lib.PRECISIONS[prec]
hXi = torch.randint(0, 4, (nh, B, 26), dtype=torch.int64).pin_memory()
hXv = torch.rand(nh, B, 13).pin_memory()
hout = torch.empty(nh, B).pin_memory()
ws = torch.zeros(lib.dfw_forward_host_stream_workspace_bytes(plan.model_ref, B, P) + 4096, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.dfw_forward_host_stream(plan.model_ref, hXi.data_ptr(), hXv.data_ptr(), nh * B, B, P, ws.data_ptr(), ws.numel(), None, hout.data_ptr(), st)
    _lib.check(rc, "x")
for _ in range(3): run()
t0 = time.perf_counter()
for _ in range(10): run()
t = (time.perf_counter() - t0) / (10 * nh)
print(f"{prec}: {t * 1e6:.1f} us/step -> {B / t / 1e6:.1f} M samples/s")
