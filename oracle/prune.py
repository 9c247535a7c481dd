"""Restatement of the reference's magnitude pruning (test infrastructure only).

* ``bisect_threshold``  -- model/DeepFMs.py:807-823 (``binary_search_threshold``):
  bisection on t in [0, 100] until ``mean(|w| < t)`` is within 1e-4 of the target,
  at most 101 probes, returns the last midpoint.
* ``one_shot_prune``    -- the pruning block of ``fit`` (model/DeepFMs.py:647-673)
  applied once at the full target rate:
  per-tensor rate ``s`` on every parameter whose name contains both ``linear`` and
  ``weight`` (the MLP layers and ``fwfm_linear``), one global threshold at rate
  ``s * emb_r`` over the concatenated ``fm_2nd_embeddings`` tensors, and a
  symmetric mask on ``field_cov.weight`` where ``|0.5 (R + R^T)| < t`` at rate
  ``s * emb_corr``.  Biases, ``net_1_fc`` and first-order tables are untouched.
The result is zeros written into dense arrays, exactly what the reference saves.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def bisect_threshold(values: np.ndarray, target: float, total: int) -> float:
    lo, hi = 0.0, 1e2
    probes = 0
    mid = 0.0
    a = np.abs(values)
    while lo < hi:
        probes += 1
        mid = (lo + hi) / 2
        rate = float((a < mid).sum()) / total
        if abs(rate - target) < 0.0001:
            return mid
        if rate > target:
            hi = mid
        else:
            lo = mid
        if probes > 100:
            break
    return mid


def one_shot_prune(params: Dict[str, np.ndarray], sparse: float = 0.9, emb_r: float = 0.444,
                   emb_corr: float = 1.0, prune_fm: bool = True, prune_deep: bool = True,
                   prune_r: bool = True) -> Dict[str, np.ndarray]:
    out = {k: np.array(v, copy=True) for k, v in params.items()}
    if prune_fm:
        names = [k for k in out if "fm_2nd_embeddings" in k]
        stacked = np.concatenate([out[k] for k in names], axis=0)
        t_emb = bisect_threshold(stacked, sparse * emb_r, stacked.size)
        for k in names:
            out[k][np.abs(out[k]) < np.float32(t_emb)] = 0
    for k in list(out):
        if "linear" in k and "weight" in k and prune_deep:
            t = bisect_threshold(out[k], sparse, out[k].size)
            out[k][np.abs(out[k]) < np.float32(t)] = 0
        if k == "field_cov.weight" and prune_r:
            sym = np.float32(0.5) * (out[k] + out[k].T)
            t = bisect_threshold(sym, sparse * emb_corr, out[k].size)
            out[k][np.abs(sym) < np.float32(t)] = 0
    return out
