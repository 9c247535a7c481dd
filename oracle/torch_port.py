"""fp32 torch-CPU restatement of ``DeepFMs.forward`` that keeps the reference's op sequence.

TEST INFRASTRUCTURE ONLY -- this is the checker and the timed CPU baseline
(``bench.py --impl reference`` / ``cpu_baseline``), never the product path.

Unlike ``closed_form`` this function issues the same aten operations, on the same
shapes, as the reference forward so that its CPU cost is representative:

* one lookup + Xv scale per field ............ model/DeepFMs.py:300-335
* ``stack`` to (F, B, K) ....................... model/DeepFMs.py:337
* two einsums for the fwlw term .............. model/DeepFMs.py:344-345
* the materialised (F, F, B, K) outer product  model/DeepFMs.py:352
* its field_cov-weighted copy ................ model/DeepFMs.py:363-364
* double sum minus the diagonal, halved ...... model/DeepFMs.py:354-355, 366-367
* ``cat`` to (B, F*K), Linear/ReLU chain, fc .. model/DeepFMs.py:398-428
* use_lw projection and the final sum ........ model/DeepFMs.py:445-469

Dropout layers are identities in ``eval()`` and are omitted.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as TF

from .config import PathConfig


def _lookup(cfg: PathConfig, sd: Dict[str, torch.Tensor], prefix: str, f: int,
            idx: torch.Tensor) -> torch.Tensor:
    """idx (B, 1) int64 -> (B, width).  Bag-of-one sum == plain row."""
    if cfg.is_qr(f):
        c = cfg.qr_collisions
        q = torch.div(idx, c, rounding_mode="floor")
        r = torch.remainder(idx, c)
        eq = TF.embedding_bag(q, sd[f"{prefix}.{f}.weight_q"], mode="sum")
        er = TF.embedding_bag(r, sd[f"{prefix}.{f}.weight_r"], mode="sum")
        if cfg.qr_operation == "mult":
            return eq * er
        if cfg.qr_operation == "add":
            return eq + er
        raise ValueError("qr_operation 'concat' is not on the hot path")
    w = sd[f"{prefix}.{f}.weight"]
    if cfg.embedding_bag:
        return TF.embedding_bag(idx, w, mode="sum")
    return TF.embedding(idx, w).sum(1)


def _field_list(cfg, sd, prefix, Xi, Xv, zero):
    out = []
    for f in range(cfg.field_size):
        if f < cfg.numerical:
            out.append((_lookup(cfg, sd, prefix, f, zero).t() * Xv[:, f]).t())
        else:
            out.append(_lookup(cfg, sd, prefix, f, Xi[:, f - cfg.numerical, :].contiguous()))
    return out


@torch.no_grad()
def forward(cfg: PathConfig, sd: Dict[str, torch.Tensor], Xi: torch.Tensor, Xv: torch.Tensor,
            return_parts: bool = False):
    """Xi (B, F-num, 1) int64, Xv (B, num) fp32 -> logits (B,) fp32."""
    if not (cfg.use_fm or cfg.use_fwfm):
        raise ValueError("hot path needs use_fm or use_fwfm")
    zero = torch.zeros(Xi.shape[0], 1, dtype=torch.long)

    if not cfg.use_fwlw:
        first = torch.cat(_field_list(cfg, sd, "fm_1st_embeddings", Xi, Xv, zero), 1)   # (B, F)
    rows = _field_list(cfg, sd, "fm_2nd_embeddings", Xi, Xv, zero)
    E = torch.stack(rows)                                                                # (F, B, K)
    if cfg.use_fwlw:
        scaled = torch.einsum("ijk,ik->ijk", E, sd["fwfm_linear.weight"])
        first = torch.einsum("ijk->ji", scaled)                                          # (B, F)

    outer = torch.einsum("kij,lij->klij", E, E)                                          # (F, F, B, K)
    if cfg.use_fwfm:
        R = sd["field_cov.weight"]
        outer = torch.einsum("klij,kl->klij", outer, (R.t() + R) * 0.5)
    second = (outer.sum(0).sum(0) - torch.einsum("kkij->kij", outer).sum(0)) * 0.5       # (B, K)

    if cfg.use_deep:
        x = torch.cat(rows, 1)                                                           # (B, F*K)
        for l in range(1, cfg.h_depth + 1):
            x = torch.relu(TF.linear(x, sd[f"net_1_linear_{l}.weight"], sd[f"net_1_linear_{l}.bias"]))
        deep = TF.linear(x, sd["net_1_fc.weight"])                                       # (B, 1)

    if cfg.use_lw:
        first = first @ sd["fm_1st.weight"].t()                                          # (B, 1)

    total = first.sum(1) + second.sum(1) + sd["bias"]
    if cfg.use_deep:
        total = first.sum(1) + second.sum(1) + deep.sum(1) + sd["bias"]
    if return_parts:
        return total, dict(E=E, first=first, second=second)
    return total
