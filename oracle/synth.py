"""Deterministic synthetic cardinalities, weights and inputs (test infrastructure only).

Weight distributions follow the reference's ``init_weights`` (model/DeepFMs.py:472-495):
first-order tables N(0,1); second-order tables N(0, 0.01^2) (optionally scaled up to a
trained-model magnitude); every ``*linear*`` tensor
N(0, glorot^2) with the glorot of the layer's weight shape shared by its bias;
``field_cov`` N(0, 1/F); ``fm_1st`` / ``net_1_fc`` N(0, 2/last_layer_size).
Values come from numpy's PCG64 so a fixture can be regenerated from (config, seed).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .config import PathConfig

# Paper Criteo cardinalities: latency/criteo_latency.cpp:38-39
CRITEO_PAPER = [1] * 13 + [1458, 556, 245197, 166166, 306, 20, 12055, 634, 4, 46330, 5229, 243454,
                           3177, 27, 11745, 225322, 11, 4727, 2058, 5, 238640, 18, 16, 67856, 89, 50942]
# feature_sizes the reference derives for the bundled tiny Criteo data
CRITEO_TINY = [1] * 13 + [1254, 521, 60857, 59255, 238, 10, 10724, 533, 3, 27022, 4679, 61746, 3045,
                          25, 8634, 65097, 9, 3798, 1772, 4, 63264, 11, 15, 31284, 51, 24958]
# Un-thresholded Kaggle display-advertising cardinalities (not in the reference)
CRITEO_KAGGLE = [1] * 13 + [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593,
                            3194, 27, 14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105,
                            142572]
# Synthetic Twitter RecSys2020 shape: 11 dense + 36 sparse (model/Datasets.py:41-42)
TWITTER_SYNTH = [1] * 11 + ([3, 3, 3, 16777216, 67, 4, 16, 16777216, 8388608, 16777216, 1048576, 2097152,
                             1048576, 4, 32, 8, 25] + [4096] * 7 + [64, 512, 128, 1048576, 1048576]
                            + [262144] * 4 + [1048576] * 3)


def make_weights(cfg: PathConfig, seed: int = 42, emb_scale: float = 10.0) -> Dict[str, np.ndarray]:
    rng = np.random.Generator(np.random.PCG64(seed))
    F, K, N = cfg.field_size, cfg.embedding_size, cfg.deep_nodes
    last = (F + K) + ((N + 1) if cfg.use_deep else 0)
    out: Dict[str, np.ndarray] = {}
    glorot = 1.0
    for name, shape in cfg.state_shapes().items():
        if name == "bias":
            w = np.full(shape, 0.01)
        elif "1st_embeddings" in name:
            w = rng.standard_normal(shape)
        elif "2nd_embeddings" in name:
            w = rng.standard_normal(shape) * (0.01 * emb_scale)
        elif "linear" in name:
            if "weight" in name:
                glorot = np.sqrt(2.0 / np.sum(shape))
            w = rng.standard_normal(shape) * glorot
        elif name == "field_cov.weight":
            w = rng.standard_normal(shape) * np.sqrt(2.0 / F / 2)
        else:  # fm_1st.weight, net_1_fc.weight
            w = rng.standard_normal(shape) * np.sqrt(2.0 / last)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def make_inputs(cfg: PathConfig, batch: int, seed: int = 0, dist: str = "uniform",
                xv: str = "int50") -> Tuple[np.ndarray, np.ndarray]:
    """Xi (B, F-num, 1) int64 and Xv (B, num) fp32.

    dist: 'uniform' over each table, or 'zipf' (alpha 1.05, clipped) -- worst case and
    realistic skew.  xv: 'int50' integer-valued in [0, 50) like the
    Criteo log-squared transform, or 'unit' U[0,1) like the MinMax-scaled Twitter columns.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    num = cfg.numerical
    cats = cfg.feature_sizes[num:]
    Xi = np.empty((batch, len(cats), 1), dtype=np.int64)
    for j, n in enumerate(cats):
        if dist == "uniform":
            Xi[:, j, 0] = rng.integers(0, n, size=batch)
        elif dist == "zipf":
            Xi[:, j, 0] = np.minimum(rng.zipf(1.05, size=batch) - 1, n - 1)
        else:
            raise ValueError(dist)
    if xv == "int50":
        Xv = rng.integers(0, 50, size=(batch, num)).astype(np.float32)
    elif xv == "unit":
        Xv = rng.random((batch, num), dtype=np.float32)
    else:
        raise ValueError(xv)
    return Xi, Xv


def weights_checksum(params: Dict[str, np.ndarray]) -> str:
    import hashlib
    h = hashlib.sha256()
    for k in sorted(params):
        h.update(k.encode())
        h.update(np.ascontiguousarray(params[k]).tobytes())
    return h.hexdigest()
