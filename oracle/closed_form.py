"""fp64 closed-form restatement of ``DeepFMs.forward`` (test infrastructure only).

Follows the reference's arithmetic, not its op sequence:

* rows / Xv scaling ........ model/DeepFMs.py:300-335 (and QR: model/QREmbeddingBag.py:156-174)
* fwlw linear term ......... model/DeepFMs.py:338-347
* FM / FwFM second order ... model/DeepFMs.py:351-367
* deep MLP ................. model/DeepFMs.py:395-436
* use_lw projection, sum ... model/DeepFMs.py:445-469

Everything is evaluated in float64 from the fp32 parameters so the result can be
used for tolerance accounting of both the reference's fp32 forward and the CUDA
kernels.  The gathered block ``E`` is additionally returned in float32, computed
with a single fp32 multiply exactly as the reference does, so that "gathered
rows are bit-exact" can be checked with ``np.array_equal``.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .config import PathConfig


def _rows(cfg: PathConfig, params: Dict[str, np.ndarray], prefix: str, f: int,
          idx: np.ndarray, dtype) -> np.ndarray:
    """Row lookup of table ``f`` for indices ``idx`` (B,) -> (B, width) in ``dtype``.

    QR tables: q = idx // c, r = idx mod c, row = Wq[q] (*|+) Wr[r]
    (model/QREmbeddingBag.py:157-172).
    """
    if cfg.is_qr(f):
        c = cfg.qr_collisions
        q, r = idx // c, idx % c
        # rows first, cast after (same values; a 16.7 M-row table is not converted to look up a few hundred rows)
        wq = params[f"{prefix}.{f}.weight_q"][q].astype(dtype)
        wr = params[f"{prefix}.{f}.weight_r"][r].astype(dtype)
        if cfg.qr_operation == "mult":
            return wq * wr
        if cfg.qr_operation == "add":
            return wq + wr
        raise ValueError("qr_operation 'concat' changes the embedding width; not on the hot path")
    return params[f"{prefix}.{f}.weight"][idx].astype(dtype)


def gather_block(cfg: PathConfig, params, Xi: np.ndarray, Xv: np.ndarray,
                 prefix: str = "fm_2nd_embeddings", dtype=np.float32) -> np.ndarray:
    """E[b, f, :] as the reference builds it (model/DeepFMs.py:312-337), (B, F, width)."""
    Xi = np.asarray(Xi).reshape(Xi.shape[0], -1)
    B, num = Xi.shape[0], cfg.numerical
    width = 1 if prefix == "fm_1st_embeddings" else cfg.embedding_size
    E = np.empty((B, cfg.field_size, width), dtype=dtype)
    zero = np.zeros(B, dtype=np.int64)
    for f in range(cfg.field_size):
        if f < num:
            # numeric: row 0 of a one-row table scaled by Xv (one multiply)
            E[:, f, :] = _rows(cfg, params, prefix, f, zero, dtype) * Xv[:, f].astype(dtype)[:, None]
        else:
            # categorical: plain row copy, Xv is NOT applied (implicit 1)
            E[:, f, :] = _rows(cfg, params, prefix, f, Xi[:, f - num].astype(np.int64), dtype)
    return E


def forward(cfg: PathConfig, params: Dict[str, np.ndarray], Xi, Xv) -> Dict[str, np.ndarray]:
    """Closed form in fp64.  Returns dict(E32, first, second, deep, logit, prob)."""
    Xi = np.asarray(Xi)
    Xv = np.asarray(Xv, dtype=np.float32)
    F = cfg.field_size
    if not (cfg.use_fm or cfg.use_fwfm):
        raise ValueError("hot path needs use_fm or use_fwfm")
    E32 = gather_block(cfg, params, Xi, Xv, dtype=np.float32)
    # fp64 view of the SAME fp32 values the model feeds downstream
    E = E32.astype(np.float64)
    B = E.shape[0]

    # ---- first order -----------------------------------------------------
    if cfg.use_fwlw:
        wl = params["fwfm_linear.weight"].astype(np.float64)            # (F, K)
        first_vec = np.einsum("bfk,fk->bf", E, wl)                       # (B, F)
    else:
        first_vec = gather_block(cfg, params, Xi, Xv, prefix="fm_1st_embeddings",
                                 dtype=np.float32).astype(np.float64)[:, :, 0]
    if cfg.use_lw:
        first = first_vec @ params["fm_1st.weight"].astype(np.float64)[0]
    else:
        first = first_vec.sum(axis=1)

    # ---- second order ----------------------------------------------------
    if cfg.use_fwfm:
        W = params["field_cov.weight"].astype(np.float32)
        # the reference symmetrises in fp32: (W.t() + W) * 0.5  (model/DeepFMs.py:364)
        Rs = ((W.T + W) * np.float32(0.5)).astype(np.float64)
    else:
        Rs = np.ones((F, F), dtype=np.float64)
    U = np.triu(Rs, k=1)                                                 # i < j only
    second = np.einsum("bik,ij,bjk->b", E, U, E)

    # ---- deep ------------------------------------------------------------
    if cfg.use_deep:
        x = E.reshape(B, F * cfg.embedding_size)
        for l in range(1, cfg.h_depth + 1):
            w = params[f"net_1_linear_{l}.weight"].astype(np.float64)
            b = params[f"net_1_linear_{l}.bias"].astype(np.float64)
            x = np.maximum(x @ w.T + b, 0.0)
        deep = x @ params["net_1_fc.weight"].astype(np.float64)[0]
    else:
        deep = np.zeros(B, dtype=np.float64)

    logit = first + second + deep + float(params["bias"][0])
    prob = 1.0 / (1.0 + np.exp(-logit))
    return dict(E32=E32, first=first, second=second, deep=deep, logit=logit, prob=prob)
