"""The oracle restatements against every golden fixture produced by the reference itself."""
import numpy as np
import pytest
import torch

from oracle import closed_form, torch_port
from golden_util import CASES, load_case, load_tiny, tiny_weights, logit_tol


def test_fixture_inventory():
    assert len(CASES) >= 20


@pytest.mark.parametrize("name", CASES)
def test_closed_form_matches_reference(name):
    c = load_case(name)
    out = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])
    # gathered block: bit-exact against what the reference's own lookups produced
    assert np.array_equal(out["E32"], c["E"])
    # fp64 closed form vs the reference's fp32 forward: fp32 rounding of the reference only
    assert np.abs(out["logit"] - c["logits"]).max() <= logit_tol(c["logits"], 1e-5)


@pytest.mark.parametrize("name", CASES)
def test_torch_port_matches_reference(name):
    c = load_case(name)
    sd = {k: torch.from_numpy(v) for k, v in c["weights"].items()}
    got = torch_port.forward(c["cfg"], sd, torch.from_numpy(c["Xi"]), torch.from_numpy(c["Xv"])).numpy()
    # same op sequence as the reference -> agreement to a few fp32 ulps of the logit scale
    assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], 2e-6)


@pytest.mark.parametrize("tag", ["lw1", "lw0", "deep_fwlw"])
def test_tiny_criteo_config1(tag):
    from sklearn.metrics import roc_auc_score, log_loss
    t = load_tiny()
    v = t["variants"][tag]
    w = tiny_weights(v)
    out = closed_form.forward(v["cfg"], w, t["Xi"], t["Xv"])
    assert np.abs(out["logit"] - v["logits"]).max() <= logit_tol(v["logits"], 1e-5)
    # the reference's callers apply an fp32 sigmoid before the metric (model/DeepFMs.py:777);
    # saturation ties are part of its AUC, so the oracle's logits go through the same fp32 sigmoid
    prob = torch.sigmoid(torch.from_numpy(out["logit"].astype(np.float32))).numpy().astype(np.float64)
    auc = roc_auc_score(t["y"], prob)
    assert abs(auc - v["metrics"][1]) <= 1e-6
    # parameter census the reference logs for tiny-criteo
    assert sum(v["cfg"].feature_sizes) == 428822
    n2 = sum(int(np.prod(a.shape)) for k, a in w.items() if "2nd_embeddings" in k)
    assert n2 == 4288220
    if tag == "lw1":
        assert sum(int(np.prod(a.shape)) for a in w.values()) == 4718564 + 39
    if tag == "lw0":
        assert sum(int(np.prod(a.shape)) for a in w.values()) == 4718564
