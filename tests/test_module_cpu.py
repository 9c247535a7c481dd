"""Host-side mirror of the reference module: names, shapes, construction parity, refusal behaviour (no GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from golden_util import CASES, GOLDEN, load_case
from oracle import synth
from xsdeepfwfm_deprecated_b200.model import DeepFMs, QREmbeddingBag

SMALL = [1] * 13 + [7, 313, 12, 1999, 3, 250, 45, 201, 2, 1024, 77, 5, 640, 9, 33, 4096, 11, 200, 58, 4,
                    900, 18, 16, 129, 89, 2500]


def build(cfg, **kw):
    return DeepFMs(cfg.field_size, cfg.feature_sizes, embedding_size=cfg.embedding_size, h_depth=cfg.h_depth,
                   deep_nodes=cfg.deep_nodes, use_fm=cfg.use_fm, use_fwfm=cfg.use_fwfm, use_deep=cfg.use_deep,
                   use_fwlw=cfg.use_fwlw, use_lw=cfg.use_lw, use_cuda=False, numerical=cfg.numerical,
                   embedding_bag=cfg.embedding_bag, qr_flag=cfg.qr_flag, qr_operation=cfg.qr_operation,
                   qr_collisions=cfg.qr_collisions, qr_threshold=cfg.qr_threshold, **kw)


@pytest.mark.parametrize("name", CASES)
def test_state_dict_names_and_strict_load(name):
    c = load_case(name)
    m = build(c["cfg"])
    sd = m.state_dict()
    assert set(sd) == set(c["weights"])          # the key set the reference itself accepted (make_golden.py)
    for k, v in sd.items():
        assert tuple(v.shape) == c["weights"][k].shape, k
    m.load_state_dict({k: torch.from_numpy(v) for k, v in c["weights"].items()}, strict=True)


def test_construction_and_init_match_the_reference_rng_for_rng():
    z = np.load(os.path.join(GOLDEN, "ctor_parity.npz"))
    for tag in ("plain", "fwlw_bag", "qr", "fm"):
        kw = json.loads(str(z[f"{tag}::kw"]))
        m = DeepFMs(39, SMALL, use_cuda=False, random_seed=42, **kw)
        names = json.loads(str(z[f"{tag}::names"]))
        assert list(m.state_dict().keys()) == list(names.keys())      # same registration order
        assert synth.weights_checksum({k: v.numpy() for k, v in m.state_dict().items()}) == str(z[f"{tag}::ctor"]), tag
        m.init_weights()
        assert synth.weights_checksum({k: v.numpy() for k, v in m.state_dict().items()}) == str(z[f"{tag}::init"]), tag


def test_no_cpu_fallback():
    c = load_case("fwfm")
    m = build(c["cfg"]).eval()
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.from_numpy(c["Xi"]), torch.from_numpy(c["Xv"]))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            DeepFMs(39, SMALL, use_fwfm=True, use_fm=False, use_cuda=True)


@pytest.mark.parametrize("kw", [dict(use_ffm=True, use_fm=False), dict(use_logit=1, use_fm=False),
                                dict(use_fm=False, use_deep=True), dict(use_fm=False, use_deep=False),
                                dict(num_deeps=2), dict(is_batch_norm=True), dict(static_quantization=True),
                                dict(qr_flag=1, qr_operation="concat"), dict(use_fm=True, use_fwfm=True),
                                dict(precision="fp8")])
def test_unsupported_switches_raise(kw):
    with pytest.raises(ValueError):
        DeepFMs(39, SMALL, use_cuda=False, **kw)


def test_fit_refuses():
    m = DeepFMs(39, SMALL, use_cuda=False)
    with pytest.raises(NotImplementedError):
        m.fit([], [], [])


def test_qr_container_shapes():
    t = QREmbeddingBag(1001, 10, 4, operation="mult", mode="sum")
    assert tuple(t.weight_q.shape) == (251, 10) and tuple(t.weight_r.shape) == (4, 10)
    assert float(t.weight_q.min()) >= np.sqrt(1 / 1001) - 1e-7      # uniform_(sqrt(1/n), 1): the reference's real init
    with pytest.raises(ValueError):
        QREmbeddingBag(10, 4, 2, operation="concat")


def test_threshold_bisection_matches_oracle():
    from oracle import prune
    rng = np.random.default_rng(0)
    w = rng.standard_normal((400, 390)).astype(np.float32) * 0.05
    m = DeepFMs(39, SMALL, use_cuda=False)
    t = m.binary_search_threshold(torch.from_numpy(w), 0.9, w.size)
    assert t == prune.bisect_threshold(w, 0.9, w.size)
    assert abs(float((np.abs(w) < t).mean()) - 0.9) < 1e-4


def test_index_dtype_switch_is_validated():
    c = load_case("deepfwfm_fwlw")
    assert build(c["cfg"]).index_dtype == "int64"              # the reference's LongTensor format is the default
    assert build(c["cfg"], index_dtype="int32").index_dtype == "int32"
    with pytest.raises(ValueError):
        build(c["cfg"], index_dtype="int16")


def test_eval_and_train_keep_the_plan():
    """predict_proba / eval_by_batch call eval() on every call; that must not throw the packed images away.
    Parameter edits are caught by stale() (data_ptr + _version), device moves and load_state_dict still drop the plan."""
    c = load_case(CASES[0])
    m = build(c["cfg"]).eval()
    sentinel = object()
    m._plan = sentinel
    m.eval(); m.eval()
    assert m._plan is sentinel
    m.train()                                    # a mode CHANGE drops it: fit() prunes with .data[mask] = 0 in train mode
    assert m._plan is None
    m._plan = sentinel
    m.train(); m.train()
    assert m._plan is sentinel
    m.eval()
    assert m._plan is None
    m._plan = sentinel
    m.load_state_dict(m.state_dict())
    assert m._plan is None
    m._plan = sentinel
    m.float()                                    # _apply: tensors may be re-bound
    assert m._plan is None
