"""Debug tool: host-side cost per call of dfw_forward (enqueue only) and of the streamed host API pieces."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = 4096
m = bench.make_model(dev, prec, synth.CRITEO_PAPER)
plan = m._get_plan(); plan.ensure_image(m, prec)
Xi, Xv = bench.make_batches(dev, synth.CRITEO_PAPER, B, 4, seed=0)
out = torch.zeros(B, device=dev)
ws = plan.get_workspace(lib.dfw_forward_workspace_bytes(plan.model_ref, B, _lib.PRECISIONS[prec]))
st = torch.cuda.current_stream().cuda_stream
def fwd():
    lib.dfw_forward(plan.model_ref, Xi[0].data_ptr(), 26, 1, Xv[0].data_ptr(), 13, 1, B, _lib.PRECISIONS[prec], ws.data_ptr(), ws.numel(), out.data_ptr(), None, None, st)
for _ in range(10): fwd()
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for _ in range(n): fwd()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"dfw_forward: host enqueue {1e6*(t1-t0)/n:.1f} us/call; with drain {1e6*(t2-t0)/n:.1f} us/call")
hXi = torch.empty(B, 26, dtype=torch.int64).pin_memory(); dXi = torch.empty(B, 26, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n): dXi.copy_(hXi, non_blocking=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"852 KB H2D memcpyAsync (torch): host enqueue {1e6*(t1-t0)/n:.1f} us/call; with drain {1e6*(t2-t0)/n:.1f} us/call")
