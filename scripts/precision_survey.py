"""Measured error of every precision on every golden case (decides the bounds written in tests/test_fullsize_gpu.py):
max|dlogit| / max|logit|, max|d deep| / max|deep|."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from golden_util import CASES, load_case
from oracle import closed_form
from test_parity_gpu import to_cuda, run

for name in CASES:
    c = load_case(name)
    if not c["cfg"].use_deep:
        continue
    ref = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])
    shallow = ref["first"] + ref["second"] + float(c["weights"]["bias"][0])
    ml, md = np.abs(ref["logit"]).max(), np.abs(ref["deep"]).max()
    row = f"{name:22s} max|logit| {ml:8.3f} max|deep| {md:7.3f}"
    for prec in ("fp32", "bf16x3", "bf16"):
        try:
            got = run(to_cuda(c["cfg"], c["weights"], precision=prec), c["Xi"], c["Xv"]).astype(np.float64)
            row += f" | {prec}: {np.abs(got - ref['logit']).max() / ml:.2e} deep {np.abs((got - shallow) - ref['deep']).max() / md:.2e}"
        except Exception as ex:
            row += f" | {prec}: {str(ex)[:30]}"
    print(row, flush=True)
