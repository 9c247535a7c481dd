"""HBM roofline of the gather + interaction kernel by itself (FwFM without the deep part, BASELINE config 1's model shape)
on tables larger than L2: un-thresholded Kaggle cardinalities, plain tables (33.8 M rows, 1.35 GB fp32), uniform indices.

    python scripts/embed_roofline.py > profiles/<round>_embed_roofline.json

Algorithmic bytes per sample: 26 x 8 (Xi) + 13 x 4 (Xv) + 26 x 40 (rows) + 4 (logit) = 1304.  A random 40-byte
row costs two 32-byte DRAM sectors, so 1304 algorithmic bytes are >= 1928 DRAM bytes: 0.68 of peak is the ceiling of `frac`.
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
from xsdeepfwfm_deprecated_b200.model import DeepFMs
import bench

lib = _lib.load()
dev = torch.device("cuda", 0)
sizes = synth.CRITEO_KAGGLE
m = DeepFMs(39, sizes, use_fm=False, use_fwfm=True, use_deep=False, use_fwlw=True, use_lw=False, use_cuda=True, random_seed=42)
m = m.to(dev)
m.init_weights()
m = m.eval().freeze()
plan = m._get_plan()
peak = bench.peaks()
out = {"workload": "FwFM (use_deep=0, fwlw) gather + Xv scale + first order + FwFM second order + logit, plain Kaggle-cardinality "
                   "tables (33.8 M rows, 1.35 GB), uniform indices", "peak_GBps": peak["hbm"], "peak_source": peak["src"], "runs": []}
for B in (4096, 65536, 524288, 2097152):
    g = torch.Generator(device=dev); g.manual_seed(B)
    cats = torch.tensor(sizes[13:], device=dev, dtype=torch.float64)
    Xi = torch.minimum((torch.rand(B, 26, generator=g, device=dev, dtype=torch.float64) * cats).long(), (cats - 1).long()).unsqueeze(-1).contiguous()
    Xv = torch.randint(0, 50, (B, 13), generator=g, device=dev).float()
    logits = torch.empty(B, device=dev)
    ws = torch.zeros(lib.dfw_forward_workspace_bytes(plan.model_ref, B, 0) + 4096, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.check(lib.dfw_forward(plan.model_ref, Xi.data_ptr(), 26, 1, Xv.data_ptr(), 13, 1, B, 0, ws.data_ptr(), ws.numel(),
                                   logits.data_ptr(), None, None, st), "dfw_forward")
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    n = 20 if B >= 65536 else 200
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        run()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / n
    with torch.no_grad():
        ok = bool(torch.equal(m(Xi[:64], Xv[:64]), logits[:64]))
    gbps = 1304 * B / (ms * 1e-3) / 1e9
    out["runs"].append(dict(B=B, ms=round(ms, 4), samples_per_s=round(B / (ms * 1e-3), 1), algorithmic_GBps=round(gbps, 1),
                            frac_of_peak=round(gbps / peak["hbm"], 4), launches_per_call=2, consistent=ok))
print(json.dumps(out, indent=1))
