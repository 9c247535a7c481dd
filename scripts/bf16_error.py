"""Report the actual error of the tensor-core (bf16) path against the fp64 oracle (used to state its bound)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import closed_form, synth
from oracle.config import PathConfig
from test_parity_gpu import to_cuda, run
from golden_util import CASES, load_case
rows = []
for name in CASES:
    c = load_case(name)
    if not c["cfg"].use_deep or c["cfg"].field_size * c["cfg"].embedding_size > 512: continue
    ref = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])
    got = run(to_cuda(c["cfg"], c["weights"], precision="bf16"), c["Xi"], c["Xv"])
    d = np.abs(got - ref["logit"]).max()
    rows.append((name, d / np.abs(ref["logit"]).max(), d / max(np.abs(ref["deep"]).max(), 1e-30), np.abs(ref["deep"]).max()))
cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
w = synth.make_weights(cfg, seed=42); Xi, Xv = synth.make_inputs(cfg, 4096, seed=0)
ref = closed_form.forward(cfg, w, Xi, Xv); got = run(to_cuda(cfg, w, precision="bf16"), Xi, Xv)
d = np.abs(got - ref["logit"]).max()
rows.append(("config2_B4096", d / np.abs(ref["logit"]).max(), d / np.abs(ref["deep"]).max(), np.abs(ref["deep"]).max()))
p32 = 1 / (1 + np.exp(-ref["logit"])); pg = 1 / (1 + np.exp(-got.astype(np.float64)))
print(f"{'case':24s} err/max|logit|  err/max|deep|  max|deep|")
for r in rows: print(f"{r[0]:24s} {r[1]:12.3e} {r[2]:12.3e} {r[3]:10.4f}")
print("config2 max |dprob|", np.abs(p32 - pg).max())
