"""A few fused forwards of one workload (for ncu / instruction counts).  python scripts/run_one.py WORKLOAD [B] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xsdeepfwfm_deprecated_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "criteo"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
bench.set_workload(wl.replace("_small", ""))
if wl.endswith("_small"):
    from xsdeepfwfm_deprecated_b200.utils import workloads
    bench.SIZES = workloads.CRITEO_PAPER
lib = _lib.load()
dev = torch.device("cuda", 0)
m = bench.make_model(dev, prec, bench.SIZES)
plan = m._get_plan(); plan.ensure_image(m, prec)
Xi, Xv = bench.make_batches(dev, bench.SIZES, B, 2, seed=0)
out = torch.zeros(B, device=dev)
for j in range(4):
    rc = lib.dfw_forward_fused(plan.model_ref, Xi[j % 2].data_ptr(), bench.CATS, 1, Xv[j % 2].data_ptr(), bench.NUM, 1, B,
                               _lib.PRECISIONS[prec], out.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "dfw_forward_fused")
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
