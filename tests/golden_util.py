"""Loading of the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
import glob
import json
import os

import numpy as np

from oracle import prune, synth
from oracle.config import PathConfig

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if os.path.basename(p) not in ("tiny_criteo.npz", "ctor_parity.npz", "loader_parity.npz"))
_cache = {}


def load_case(name):
    if name in _cache:
        return _cache[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = PathConfig.from_json(json.loads(str(z["cfg"])))
    w = synth.make_weights(cfg, seed=int(z["seed"]), emb_scale=float(z["emb_scale"]))
    if int(z["pruned"]):
        w = prune.one_shot_prune(w, 0.9, 0.444, 1.0)
    assert synth.weights_checksum(w) == str(z["checksum"]), f"{name}: regenerated weights drifted"
    stored = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    for k, v in stored.items():
        assert np.array_equal(v, w[k]), f"{name}: stored weight {k} differs from regenerated"
    Xi = z["Xi"].astype(np.int64)
    out = dict(cfg=cfg, weights=w, Xi=Xi, Xv=z["Xv"].astype(np.float32), logits=z["logits"], E=z["E"])
    _cache[name] = out
    return out


def load_tiny():
    if "tiny" in _cache:
        return _cache["tiny"]
    z = np.load(os.path.join(GOLDEN, "tiny_criteo.npz"))
    out = dict(y=z["y"].astype(np.float64), Xv=z["Xv"].astype(np.float32),
               Xi=z["Xi"].astype(np.int64).reshape(-1, 26, 1), variants={})
    for tag in ("lw1", "lw0", "deep_fwlw"):
        cfg = PathConfig.from_json(json.loads(str(z[f"{tag}::cfg"])))
        out["variants"][tag] = dict(cfg=cfg, logits=z[f"{tag}::logits"], metrics=z[f"{tag}::metrics"],
                                    checksum=str(z[f"{tag}::checksum"]), seed=int(z["seed"]),
                                    emb_scale=float(z["emb_scale"]))
    _cache["tiny"] = out
    return out


def tiny_weights(variant):
    w = synth.make_weights(variant["cfg"], seed=variant["seed"], emb_scale=variant["emb_scale"])
    assert synth.weights_checksum(w) == variant["checksum"]
    return w


def logit_tol(ref, rel=1e-5):
    """fp32 parity bound: |delta| <= rel * max|logit_ref|."""
    return rel * float(np.abs(ref).max())
