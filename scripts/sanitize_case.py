"""One fused forward at B = 4096 and at a ragged B, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_case.py [precision] [B ...]

Criteo shape (F = 39, K = 10, MLP 400x400x400) with small tables so the run is about the kernels, not the set-up.  Checks the
result against the oracle as well, so a sanitizer-clean but wrong run cannot pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import closed_form, synth
from oracle.config import PathConfig
from test_parity_gpu import to_cuda, run

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
Bs = [int(a) for a in sys.argv[2:]] or [4096, 333]
sizes = [1] * 13 + [7, 313, 12, 1999, 3, 250, 45, 201, 2, 1024, 77, 5, 640, 9, 33, 4096, 11, 200, 58, 4, 900, 18, 16, 129, 89, 2500]
cfg = PathConfig(39, sizes, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
w = synth.make_weights(cfg, seed=7)
m = to_cuda(cfg, w, precision=prec)
for B in Bs:
    Xi, Xv = synth.make_inputs(cfg, B, seed=B)
    ref = closed_form.forward(cfg, w, Xi, Xv)["logit"]
    got = run(m, Xi, Xv)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"sanitize_case {prec} B={B}: max|dlogit|/max|logit| = {err:.2e}", flush=True)
    assert err <= (5e-4 if prec == "bf16" else 1e-5)
print("sanitize_case: ok")
