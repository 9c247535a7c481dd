"""One-shot pruning of BASELINE config 3 on the device vs the reference recipe.

    python scripts/prune_bench.py > profiles/<round>_prune_bench.json

device      DeepFMs.prune_one_shot: per tensor set one cooperative bisection launch + one masking launch (CUDA events)
torch_gpu   the reference's own loop (binary_search_threshold: one (abs(p) < mid).sum().item() per probe) on the same GPU tensors
cpu_port    oracle/prune.py (numpy restatement) on the host
bytes       algorithmic: 4 B per element per probe for every bisection + 8 B per element (read + write) for every mask
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import prune, synth
from oracle.config import PathConfig
from xsdeepfwfm_deprecated_b200.model import DeepFMs

cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
w = synth.make_weights(cfg, seed=42)
sd = {k: torch.from_numpy(v) for k, v in w.items()}


def fresh():
    m = DeepFMs(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True, use_cuda=True)
    m.load_state_dict(sd)
    return m.cuda().eval()


def reference_loop_on_gpu(m, sparse, emb_r, emb_corr):
    """model/DeepFMs.py:647-673 with its own binary_search_threshold, on CUDA tensors through torch ops."""
    def bst(param, target, total):
        l, r, cnt, mid = 0., 1e2, 0, 0.
        while l < r:
            cnt += 1
            mid = (l + r) / 2
            rate = (abs(param) < mid).sum().item() * 1.0 / total
            if abs(rate - target) < 0.0001:
                return mid
            elif rate > target:
                r = mid
            else:
                l = mid
            if cnt > 100:
                break
        return mid
    stacked = torch.cat([p.data for n, p in m.named_parameters() if "fm_2nd_embeddings" in n], 0)
    t_emb = bst(stacked, sparse * emb_r, stacked.numel())
    for n, p in m.named_parameters():
        if "fm_2nd_embeddings" in n:
            p.data[abs(p.data) < t_emb] = 0
        if "linear" in n and "weight" in n:
            t = bst(p.data, sparse, p.numel())
            p.data[abs(p.data) < t] = 0
        if n == "field_cov.weight":
            sym = 0.5 * (p.data + p.data.t())
            t = bst(sym, sparse * emb_corr, p.numel())
            p.data[abs(sym) < t] = 0


out = {"workload": "BASELINE config 3: one-shot prune of the config-2 model (13.7 M parameters), sparse 0.9, emb_r 0.444, emb_corr 1"}
# device
times, rep = [], None
for it in range(4):
    m = fresh()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rep = m.prune_one_shot(0.9, 0.444, 1.0)      # the report read at the end is the only host synchronisation
    e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
dev_state = {k: v.detach().cpu() for k, v in m.state_dict().items()}
n_el = {"emb": sum(v.size for k, v in w.items() if "fm_2nd_embeddings" in k)}
n_el.update({k: w[k].size for k in rep if k != "emb"})
alg = sum(4 * n_el[k] * rep[k][1] + 8 * n_el[k] for k in rep)
best = min(times[1:])
out["device"] = {"ms": round(best, 4), "launches": 2 * len(rep), "probes": {k: v[1] for k, v in rep.items()},
                 "algorithmic_bytes": alg, "GBps": round(alg / (best * 1e-3) / 1e9, 1),
                 "host_synchronisations": 1}
# reference loop on the GPU
times = []
for it in range(3):
    m2 = fresh()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reference_loop_on_gpu(m2, 0.9, 0.444, 1.0)
    torch.cuda.synchronize()
    times.append((time.perf_counter() - t0) * 1e3)
same = all(torch.equal(dev_state[k], v.detach().cpu()) for k, v in m2.state_dict().items())
out["torch_gpu_reference_loop"] = {"ms": round(min(times[1:]), 3), "bit_identical_to_device": same,
                                   "host_synchronisations": sum(v[1] for v in rep.values())}
# CPU port
t0 = time.perf_counter()
want = prune.one_shot_prune(w, 0.9, 0.444, 1.0)
out["cpu_port"] = {"ms": round((time.perf_counter() - t0) * 1e3, 1), "cores": 1,
                   "bit_identical_to_device": all(np.array_equal(dev_state[k].numpy(), v) for k, v in want.items())}
out["nonzero_parameters"] = int(sum(int((v != 0).sum()) for v in dev_state.values()))
print(json.dumps(out, indent=1))
