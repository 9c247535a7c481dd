"""A few fused forwards of one large batch (for ncu captures of the persistent pipeline in steady state).
    python scripts/run_big_batch.py [B] [precision] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from xsdeepfwfm_deprecated_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lib = _lib.load()
dev = torch.device("cuda", 0)
m = bench.make_model(dev, prec, bench.SIZES)
plan = m._get_plan(); plan.ensure_image(m, prec)
Xi, Xv = bench.make_batches(dev, bench.SIZES, B, 2, seed=0)
out = torch.zeros(B, device=dev)
for j in range(reps):
    rc = lib.dfw_forward_fused(plan.model_ref, Xi[j % 2].data_ptr(), bench.CATS, 1, Xv[j % 2].data_ptr(), bench.NUM, 1, B,
                               _lib.PRECISIONS[prec], out.data_ptr(), None, None, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "dfw_forward_fused")
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
