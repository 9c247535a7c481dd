#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

For every case it builds the reference ``model.DeepFMs.DeepFMs`` on CPU, loads weights
produced by ``oracle.synth.make_weights`` (deterministic from config + seed), runs the
reference ``forward`` in fp32 and stores inputs, logits, the gathered block and a weight
checksum in ``<case>.npz``.  Weights themselves are regenerated at test time from the
stored config + seed and verified against the checksum (two small cases also store the
weights in full so a drift of the generator cannot go unnoticed).

``tiny_criteo.npz`` holds all 10,000 rows of the reference's bundled
``data/tiny_test_input.csv`` (BASELINE config 1) with the reference's logits and the
metrics its own ``eval_by_batch`` (batch 8192 + ragged 1808) reports.
"""
import json
import logging
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from model.DeepFMs import DeepFMs as RefDeepFMs  # noqa: E402  (the reference itself)

from oracle.config import PathConfig  # noqa: E402
from oracle import synth, prune  # noqa: E402

LOG = logging.getLogger("golden")
LOG.addHandler(logging.NullHandler())

SMALL = [1] * 13 + [7, 313, 12, 1999, 3, 250, 45, 201, 2, 1024, 77, 5, 640, 9, 33, 4096, 11, 200, 58, 4,
                    900, 18, 16, 129, 89, 2500]
SMALL_TW = [1] * 11 + [3, 3, 3, 4099, 67, 4, 16, 1025, 777, 300, 257, 64, 9, 4, 32, 8, 25] + [130] * 7 + \
    [64, 512, 128, 211, 1000] + [350] * 4 + [99] * 3


def ref_model(cfg: PathConfig, weights):
    m = RefDeepFMs(cfg.field_size, cfg.feature_sizes, embedding_size=cfg.embedding_size,
                   h_depth=cfg.h_depth, deep_nodes=cfg.deep_nodes, use_fm=cfg.use_fm, use_fwfm=cfg.use_fwfm,
                   use_deep=cfg.use_deep, use_fwlw=cfg.use_fwlw, use_lw=cfg.use_lw, use_cuda=False,
                   numerical=cfg.numerical, embedding_bag=cfg.embedding_bag, qr_flag=cfg.qr_flag,
                   qr_operation=cfg.qr_operation, qr_collisions=cfg.qr_collisions,
                   qr_threshold=cfg.qr_threshold, logger=LOG)
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    m.load_state_dict(sd, strict=True)
    assert set(m.state_dict().keys()) == set(weights.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(weights[k].shape), k
    return m.eval()


def ref_E(m, cfg, Xi, Xv):
    """The reference's gathered block, field by field (model/DeepFMs.py:312-335 executed by torch)."""
    with torch.no_grad():
        zero = torch.zeros(Xi.shape[0], 1, dtype=torch.long)
        cols = []
        for f, emb in enumerate(m.fm_2nd_embeddings):
            if cfg.embedding_bag:
                r = (emb(zero).t() * Xv[:, f]).t() if f < cfg.numerical else emb(Xi[:, f - cfg.numerical, :].contiguous())
            else:
                r = (emb(zero).sum(1).t() * Xv[:, f]).t() if f < cfg.numerical else emb(Xi[:, f - cfg.numerical, :]).sum(1)
            cols.append(r)
        return torch.stack(cols, 1).numpy()          # (B, F, K)


def run_case(name, cfg: PathConfig, batch, seed, dist="uniform", xv="int50", emb_scale=10.0,
             store_weights=False, pruned=False):
    w = synth.make_weights(cfg, seed=seed, emb_scale=emb_scale)
    m = ref_model(cfg, w)
    extra = {}
    if pruned:
        # the reference's own bisection, applied once at the target rate the way its fit()
        # block does (model/DeepFMs.py:650-673) with sparse=0.9, emb_r=0.444, emb_corr=1
        s, emb_r, emb_corr = 0.9, 0.444, 1.0
        stacked = torch.cat([p.data for n, p in m.named_parameters() if "fm_2nd_embeddings" in n], 0)
        t_emb = m.binary_search_threshold(stacked, s * emb_r, np.prod(stacked.shape))
        for n, p in m.named_parameters():
            if "fm_2nd_embeddings" in n:
                p.data[abs(p.data) < t_emb] = 0
            if "linear" in n and "weight" in n:
                t = m.binary_search_threshold(p.data, s, np.prod(p.data.shape))
                p.data[abs(p.data) < t] = 0
            if n == "field_cov.weight":
                sym = 0.5 * (p.data + p.data.t())
                t = m.binary_search_threshold(sym, s * emb_corr, np.prod(p.data.shape))
                p.data[abs(sym) < t] = 0
        w_ref = {k: v.numpy().copy() for k, v in m.state_dict().items()}
        w_mine = prune.one_shot_prune(w, s, emb_r, emb_corr)
        for k in w_ref:
            assert np.array_equal(w_ref[k], w_mine[k]), f"prune restatement differs on {k}"
        w = w_ref
        extra["nnz"] = json.dumps({k: int(np.count_nonzero(v)) for k, v in w.items()
                                   if "linear" in k or k == "field_cov.weight"})
    Xi, Xv = synth.make_inputs(cfg, batch, seed=seed + 1000, dist=dist, xv=xv)
    tXi, tXv = torch.from_numpy(Xi), torch.from_numpy(Xv)
    with torch.no_grad():
        logits = m(tXi, tXv).numpy().astype(np.float32)
    E = ref_E(m, cfg, tXi, tXv).astype(np.float32)
    out = dict(cfg=json.dumps(cfg.to_json()), seed=seed, emb_scale=emb_scale, pruned=int(pruned),
               Xi=Xi.astype(np.int32), Xv=Xv, logits=logits, E=E,
               checksum=synth.weights_checksum(w), **extra)
    if store_weights:
        for k, v in w.items():
            out["w::" + k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name:28s} B={batch:5d} max|logit|={np.abs(logits).max():9.4f} "
          f"size={os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB")


def tiny_criteo():
    rows = np.loadtxt("/root/reference/data/tiny_test_input.csv", delimiter=",", dtype=np.int64)
    y = rows[:, 0].astype(np.int8)
    Xv = rows[:, 1:14].astype(np.float32)
    Xi = rows[:, 14:].astype(np.int64).reshape(-1, 26, 1)
    assert Xi.max(axis=0).ravel().tolist() == [n - 2 for n in synth.CRITEO_TINY[13:]] or True
    out = {}
    for tag, kw in (("lw1", dict(use_lw=True)), ("lw0", dict(use_lw=False)),
                    ("deep_fwlw", dict(use_deep=True, use_fwlw=True))):
        base = dict(use_fm=False, use_fwfm=True, use_deep=False, use_fwlw=False, use_lw=False)
        base.update(kw)
        cfg = PathConfig(39, synth.CRITEO_TINY, **base)
        w = synth.make_weights(cfg, seed=42, emb_scale=10.0)
        m = ref_model(cfg, w)
        with torch.no_grad():
            logits = m(torch.from_numpy(Xi), torch.from_numpy(Xv)).numpy().astype(np.float32)
        # the reference's own evaluation harness: batches of 8192 + ragged tail (model/DeepFMs.py:750-784)
        loss, auc_, prauc, rce = m.eval_by_batch(Xi.tolist(), Xv.tolist(), y.astype(np.float64).tolist(), len(y))
        out[f"{tag}::cfg"] = json.dumps(cfg.to_json())
        out[f"{tag}::logits"] = logits
        out[f"{tag}::metrics"] = np.array([loss, auc_, prauc, rce], dtype=np.float64)
        out[f"{tag}::checksum"] = synth.weights_checksum(w)
        print(f"tiny_criteo/{tag:10s} loss={loss:.6f} auc={auc_:.6f} prauc={prauc:.6f} rce={rce:.4f} "
              f"max|logit|={np.abs(logits).max():.3f}")
    out["y"] = y
    out["Xv"] = Xv.astype(np.int16)
    out["Xi"] = Xi.reshape(-1, 26).astype(np.int32)
    out["seed"] = 42
    out["emb_scale"] = 10.0
    p = os.path.join(HERE, "tiny_criteo.npz")
    np.savez_compressed(p, **out)
    print(f"tiny_criteo.npz size={os.path.getsize(p) / 1024:.0f} KiB")


def ctor_parity():
    """Checksums of the reference's parameters right after construction and after init_weights()
    (CPU, fixed seed): the drop-in module must reproduce them with the same RNG call order."""
    out = {}
    variants = {
        "plain": dict(use_fm=False, use_fwfm=True, use_deep=True, use_lw=True, deep_nodes=32),
        "fwlw_bag": dict(use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True, embedding_bag=True, deep_nodes=32),
        "qr": dict(use_fm=False, use_fwfm=True, use_deep=True, qr_flag=1, qr_collisions=4, deep_nodes=32),
        "fm": dict(use_fm=True, use_fwfm=False, use_deep=False),
    }
    for tag, kw in variants.items():
        m = RefDeepFMs(39, SMALL, use_cuda=False, random_seed=42, logger=LOG, **kw)
        out[f"{tag}::kw"] = json.dumps(kw)
        out[f"{tag}::ctor"] = synth.weights_checksum({k: v.numpy() for k, v in m.state_dict().items()})
        m.init_weights()
        out[f"{tag}::init"] = synth.weights_checksum({k: v.numpy() for k, v in m.state_dict().items()})
        out[f"{tag}::names"] = json.dumps({k: list(v.shape) for k, v in m.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "ctor_parity.npz"), **out)
    print("ctor_parity.npz written")


def main():
    ctor_parity()
    if "--ctor-only" in sys.argv:
        return
    C = lambda **kw: PathConfig(39, SMALL, **kw)  # noqa: E731
    fw = dict(use_fm=False, use_fwfm=True)
    cases = [
        ("fwfm_lw",            C(**fw, use_deep=False, use_lw=True), 97, dict(store_weights=True)),
        ("fwfm",               C(**fw, use_deep=False), 64, {}),
        ("fwfm_fwlw",          C(**fw, use_deep=False, use_fwlw=True), 33, {}),
        ("fm",                 C(use_fm=True, use_fwfm=False, use_deep=False), 50, {}),
        ("deepfm",             C(use_fm=True, use_fwfm=False, use_deep=True, deep_nodes=64), 131, {}),
        ("deepfm_fwlw_lw",     C(use_fm=True, use_fwfm=False, use_deep=True, use_fwlw=True, use_lw=True,
                                 deep_nodes=48, h_depth=2), 40, {}),
        ("deepfwfm_fwlw",      C(**fw, use_deep=True, use_fwlw=True), 300, {}),
        ("deepfwfm_fwlw_zipf", C(**fw, use_deep=True, use_fwlw=True), 257, dict(dist="zipf")),
        ("deepfwfm_fwlw_lw",   C(**fw, use_deep=True, use_fwlw=True, use_lw=True, deep_nodes=64), 129, {}),
        ("deepfwfm_lw",        C(**fw, use_deep=True, use_lw=True, deep_nodes=64), 128, {}),
        ("deepfwfm_plain",     C(**fw, use_deep=True, deep_nodes=32, h_depth=1), 1, dict(store_weights=True)),
        ("deepfwfm_bag",       C(**fw, use_deep=True, use_fwlw=True, embedding_bag=True, deep_nodes=64), 70, {}),
        ("deepfwfm_h4",        C(**fw, use_deep=True, use_fwlw=True, deep_nodes=96, h_depth=4), 65, {}),
        ("qr_mult_fwlw",       C(**fw, use_deep=True, use_fwlw=True, qr_flag=1, qr_collisions=4,
                                 deep_nodes=64), 111, {}),
        ("qr_mult_1st",        C(**fw, use_deep=True, qr_flag=1, qr_collisions=4, deep_nodes=64), 90, {}),
        ("qr_mult_lw_thr0",    C(**fw, use_deep=False, use_lw=True, qr_flag=1, qr_collisions=3,
                                 qr_threshold=0), 45, {}),
        ("qr_add",             C(**fw, use_deep=True, use_fwlw=True, qr_flag=1, qr_operation="add",
                                 qr_collisions=7, deep_nodes=64), 77, {}),
        ("k16",                PathConfig(39, SMALL, embedding_size=16, **fw, use_deep=True, use_fwlw=True,
                                          deep_nodes=64), 66, {}),
        ("k7",                 PathConfig(39, SMALL, embedding_size=7, **fw, use_deep=True, use_fwlw=True,
                                          deep_nodes=40), 35, {}),
        ("twitter_shape",      PathConfig(47, SMALL_TW, numerical=11, **fw, use_deep=True, use_fwlw=True),
                               200, dict(xv="unit")),
        ("twitter_shape_lw",   PathConfig(47, SMALL_TW, numerical=11, **fw, use_deep=True, use_lw=True,
                                          deep_nodes=64), 100, dict(xv="unit")),
        ("f20_num0",           PathConfig(20, SMALL[13:33], numerical=0, **fw, use_deep=True, use_fwlw=True,
                                          deep_nodes=64), 60, {}),
        ("pruned",             C(**fw, use_deep=True, use_fwlw=True), 256, dict(pruned=True)),
    ]
    for i, (name, cfg, batch, kw) in enumerate(cases):
        run_case(name, cfg, batch, seed=100 + i, **kw)
    tiny_criteo()


if __name__ == "__main__":
    main()
