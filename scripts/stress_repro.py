"""Debug tool: run the fused forward many times on the same inputs and count runs whose bits differ from the first
(a race shows up as a rare mismatch).  Usage: python scripts/stress_repro.py [iterations]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth
from oracle.config import PathConfig
from xsdeepfwfm_deprecated_b200.model import DeepFMs
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
w = synth.make_weights(cfg, seed=42)
for precision in ("bf16x3", "bf16"):
    m = DeepFMs(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True, use_cuda=True,
                precision=precision)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    m = m.cuda().eval().freeze()
    for B in (1, 7, 32, 33, 100, 4096):
        Xi, Xv = synth.make_inputs(cfg, B, seed=B)
        xi, xv = torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda()
        # a different-size forward in between changes which kernel / how many CTAs ran just before
        Xi2, Xv2 = synth.make_inputs(cfg, 777, seed=5)
        xi2, xv2 = torch.from_numpy(Xi2).cuda(), torch.from_numpy(Xv2).cuda()
        with torch.no_grad():
            first = m(xi, xv).clone()
            outs = []
            for i in range(iters):
                if i % 3 == 0:
                    m(xi2, xv2)
                outs.append(m(xi, xv))
            torch.cuda.synchronize()
        bad = sum(0 if torch.equal(o, first) else 1 for o in outs)
        worst = max(float((o - first).abs().max()) for o in outs)
        print(f"{precision} B={B}: {bad} of {iters} runs differ from the first (max |diff| {worst:.3g})", flush=True)
