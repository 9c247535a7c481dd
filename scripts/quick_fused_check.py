"""Quick correctness + timing probe of the fused kernel (debug tool): python scripts/quick_fused_check.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import closed_form, synth
from oracle.config import PathConfig
from test_parity_gpu import to_cuda, run
cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
w = synth.make_weights(cfg, seed=42)
for B in (32, 100, 4096):
    Xi, Xv = synth.make_inputs(cfg, B, seed=0)
    ref = closed_form.forward(cfg, w, Xi, Xv)
    for prec in ("bf16", "bf16x3", "fp32"):
        m = to_cuda(cfg, w, precision=prec)
        got = run(m, Xi, Xv)
        d = np.abs(got - ref["logit"]).max() / np.abs(ref["logit"]).max()
        xi, xv = torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda()
        with torch.no_grad():
            for _ in range(5): m(xi, xv)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50): m(xi, xv)
            b.record(); torch.cuda.synchronize()
        print(f"B={B:5d} {prec:7s} err/max|logit| = {d:.3e}   {a.elapsed_time(b)/50*1e3:8.1f} us/forward (python loop)", flush=True)
