"""Parity of the CUDA path (through the C ABI) with the golden fixtures and the oracle.  Needs a B200.

Bars: gathered rows bit-exact; fp32 logits within 1e-5 * max|logit_ref|; the pruned
CSR path under the same fp32 bound; the bf16 tensor path within 5e-4 * max|logit_ref|.
"""
import numpy as np
import pytest
import torch

from golden_util import CASES, load_case, load_tiny, tiny_weights, logit_tol
from oracle import closed_form, prune, synth
from oracle.config import PathConfig
from test_module_cpu import build

pytestmark = pytest.mark.gpu
FP32_REL = 1e-5
BF16_REL = 5e-4           # bf16 path, logit-relative: holds wherever the shallow (fp32) term sets the logit scale (configs 1-4)
BF16_DEEP_REL = 2.5e-2    # bf16 path in general: the whole error is the MLP's operand rounding, <= 2.5e-2 * max|deep| (measured: <= 9.1e-3 on
                          # the golden cases, 1.33e-2 / 1.74e-2 on 136- / 512-wide two-layer MLPs whose deep term is only ~0.5)


def to_cuda(cfg, weights, **kw):
    m = build(cfg, **kw)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    return m.cuda().eval()


def run(m, Xi, Xv):
    with torch.no_grad():
        return m(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda()).cpu().numpy()


def test_library_loaded_and_device_ok():
    from xsdeepfwfm_deprecated_b200 import _lib
    _lib.require_device(0)
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("name", CASES)
def test_golden_logits_fp32(name):
    c = load_case(name)
    m = to_cuda(c["cfg"], c["weights"])
    got = run(m, c["Xi"], c["Xv"])
    assert got.shape == c["logits"].shape and got.dtype == np.float32
    assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)
    ref64 = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])["logit"]
    assert np.abs(got - ref64).max() <= logit_tol(ref64, FP32_REL)


@pytest.mark.parametrize("name", CASES)
def test_gathered_block_bit_exact(name):
    c = load_case(name)
    m = to_cuda(c["cfg"], c["weights"])
    E, shallow = m.gathered_block(torch.from_numpy(c["Xi"]).cuda(), torch.from_numpy(c["Xv"]).cuda())
    assert np.array_equal(E.cpu().numpy(), c["E"])          # rows / Xv products exactly as the reference's
    o = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])
    want = o["first"] + o["second"] + float(c["weights"]["bias"][0])
    assert np.abs(shallow.cpu().numpy() - want).max() <= logit_tol(o["logit"], FP32_REL)


@pytest.mark.parametrize("name", [n for n in CASES if load_case(n)["cfg"].use_deep])
def test_golden_logits_csr(name):
    c = load_case(name)
    m = to_cuda(c["cfg"], c["weights"], precision="fp32_csr")
    got = run(m, c["Xi"], c["Xv"])
    assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)


@pytest.mark.parametrize("name", [n for n in CASES if load_case(n)["cfg"].use_deep])
def test_golden_logits_bf16(name):
    c = load_case(name)
    m = to_cuda(c["cfg"], c["weights"], precision="bf16")
    if c["cfg"].field_size * c["cfg"].embedding_size > 512:
        # outside the tensor-core kernels' shapes (F*K <= 512): bf16 runs the CUDA-core fp32 MLP like bf16x3 does
        got = run(m, c["Xi"], c["Xv"])
        assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)
        return
    got = run(m, c["Xi"], c["Xv"])
    # the shallow part stays fp32; only the deep term carries bf16 operand rounding
    # the shallow part stays fp32: all of the error is the MLP's bf16 operand rounding, bounded against the deep term's own scale
    assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL) + BF16_DEEP_REL * np.abs(
        closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])["deep"]).max()


def test_sigmoid_epilogue_and_prob():
    c = load_case("deepfwfm_fwlw")
    m = to_cuda(c["cfg"], c["weights"])
    with torch.no_grad():
        logits, prob = m(torch.from_numpy(c["Xi"]).cuda(), torch.from_numpy(c["Xv"]).cuda(), return_prob=True)
    want = torch.sigmoid(logits)
    assert torch.allclose(prob, want, atol=2e-7, rtol=1e-6)


@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 63, 64, 65, 127, 129, 1000])
def test_ragged_batches(B):
    c = load_case("deepfwfm_fwlw")
    cfg = c["cfg"]
    Xi, Xv = synth.make_inputs(cfg, B, seed=B)
    m = to_cuda(cfg, c["weights"])
    got = run(m, Xi, Xv)
    ref = closed_form.forward(cfg, c["weights"], Xi, Xv)["logit"]
    assert got.shape == (B,)
    assert np.abs(got - ref).max() <= logit_tol(ref, FP32_REL)


def test_empty_batch():
    c = load_case("deepfwfm_fwlw")
    m = to_cuda(c["cfg"], c["weights"])
    out = m(torch.empty(0, 26, 1, dtype=torch.int64, device="cuda"), torch.empty(0, 13, device="cuda"))
    assert out.shape == (0,)


def test_non_contiguous_views():
    c = load_case("deepfwfm_fwlw")
    m = to_cuda(c["cfg"], c["weights"])
    Xi = torch.from_numpy(c["Xi"]).cuda()
    Xv = torch.from_numpy(c["Xv"]).cuda()
    B = Xi.shape[0]
    big_i = torch.zeros(B, 2 * 26, 2, dtype=torch.int64, device="cuda")
    big_i[:, ::2, :1] = Xi
    big_v = torch.zeros(2 * B, 2 * 13, device="cuda")
    big_v[::2, ::2] = Xv
    vi, vv = big_i[:, ::2, :1], big_v[::2, ::2]
    assert not vi.is_contiguous() and not vv.is_contiguous()
    with torch.no_grad():
        a, b = m(Xi, Xv), m(vi, vv)
    assert torch.equal(a, b)


def test_index_out_of_range_is_a_defined_error():
    c = load_case("fwfm")
    m = to_cuda(c["cfg"], c["weights"], check_index=True)
    Xi = torch.from_numpy(c["Xi"]).cuda().clone()
    Xi[3, 5, 0] = c["cfg"].feature_sizes[13 + 5]          # one past the end
    with pytest.raises(IndexError, match="field 18"):
        m(Xi, torch.from_numpy(c["Xv"]).cuda())
    Xi[3, 5, 0] = -1
    with pytest.raises(IndexError):
        m(Xi, torch.from_numpy(c["Xv"]).cuda())


def test_in_place_pruning_is_seen_after_eval():
    """The reference prunes with param.data[mask] = 0 inside fit() (model/DeepFMs.py:660-673) and every inference
    entry point then calls eval() (model/DeepFMs.py:757): tables and MLP weights are read live, the shallow image
    (field_cov / fwfm_linear) is rebuilt by eval()."""
    c = load_case("deepfwfm_fwlw")
    m = to_cuda(c["cfg"], c["weights"])
    before = run(m, c["Xi"], c["Xv"])
    plan = m._plan
    m.eval()                                     # already in eval mode: the packed images survive
    assert m._plan is plan
    pruned = prune.one_shot_prune(c["weights"], 0.9, 0.444, 1.0)
    m.train()                                    # fit() runs in train mode
    with torch.no_grad():
        for k, p in m.named_parameters():
            p.data[torch.from_numpy(pruned[k] == 0).cuda() & (p.data != 0)] = 0
    m.eval()
    after = run(m, c["Xi"], c["Xv"])
    ref = closed_form.forward(c["cfg"], pruned, c["Xi"], c["Xv"])["logit"]
    assert np.abs(after - ref).max() <= logit_tol(ref, FP32_REL)
    assert np.abs(after - before).max() > 1e-3
    # the same edit without leaving eval mode is picked up by repack()
    m2 = to_cuda(c["cfg"], c["weights"])
    run(m2, c["Xi"], c["Xv"])
    with torch.no_grad():
        for k, p in m2.named_parameters():
            p.data[torch.from_numpy(pruned[k] == 0).cuda() & (p.data != 0)] = 0
    m2.repack()
    again = run(m2, c["Xi"], c["Xv"])
    assert np.array_equal(again, after)


def test_table_and_mlp_edits_are_live_without_any_repack():
    c = load_case("deepfwfm_fwlw")
    m = to_cuda(c["cfg"], c["weights"])
    run(m, c["Xi"], c["Xv"])
    w = {k: v.copy() for k, v in c["weights"].items()}
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "fm_2nd_embeddings" in k or "net_1_linear" in k:
                w[k] = (w[k] * np.float32(0.5)).astype(np.float32)
                p.data.mul_(0.5)
    got = run(m, c["Xi"], c["Xv"])
    ref = closed_form.forward(c["cfg"], w, c["Xi"], c["Xv"])["logit"]
    assert np.abs(got - ref).max() <= logit_tol(ref, FP32_REL)


def test_pruned_pair_list_and_csr_agree_with_dense():
    c = load_case("pruned")
    dense = run(to_cuda(c["cfg"], c["weights"]), c["Xi"], c["Xv"])
    csr = run(to_cuda(c["cfg"], c["weights"], precision="fp32_csr"), c["Xi"], c["Xv"])
    assert np.abs(dense - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)
    assert np.abs(csr - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)
    # R given by the reference's own pruned-R dump pattern: half the entries exactly zero
    W = c["weights"]["field_cov.weight"]
    live = np.count_nonzero(np.triu(0.5 * (W + W.T), 1))
    assert live < 741 // 6          # this fixture exercises the compacted pair-list walk


def test_init_weights_rebinding_is_noticed():
    c = load_case("deepfwfm_lw")
    m = to_cuda(c["cfg"], c["weights"])
    a = run(m, c["Xi"], c["Xv"])
    torch.manual_seed(5)
    m.init_weights()
    m.eval()
    b = run(m, c["Xi"], c["Xv"])
    w = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    ref = closed_form.forward(c["cfg"], w, c["Xi"], c["Xv"])["logit"]
    assert np.abs(b - ref).max() <= logit_tol(ref, FP32_REL)
    assert np.abs(a - b).max() > 1e-3


@pytest.mark.parametrize("tag", ["lw1", "lw0", "deep_fwlw"])
def test_config1_tiny_criteo_eval_by_batch(tag):
    """BASELINE config 1: all 10,000 bundled rows through eval_by_batch (8192 + ragged 1808)."""
    t = load_tiny()
    v = t["variants"][tag]
    m = to_cuda(v["cfg"], tiny_weights(v))
    loss, auc, prauc, rce = m.eval_by_batch(t["Xi"], t["Xv"], t["y"], len(t["y"]))
    ref_loss, ref_auc, ref_prauc, ref_rce = v["metrics"]
    assert abs(auc - ref_auc) <= 1e-6
    assert abs(prauc - ref_prauc) <= 1e-5
    assert abs(loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    got = m.predict_proba_host(t["Xi"], t["Xv"], want_logits=True)[1]
    assert np.abs(got - v["logits"]).max() <= logit_tol(v["logits"], FP32_REL)


def test_config2_full_size_batch_4096():
    """BASELINE config 2 at full size: paper-Criteo cardinalities, B = 4096, dense DeepFwFM + fwlw."""
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = synth.make_weights(cfg, seed=42)
    Xi, Xv = synth.make_inputs(cfg, 4096, seed=0)
    ref = closed_form.forward(cfg, w, Xi, Xv)
    m = to_cuda(cfg, w)
    got = run(m, Xi, Xv)
    assert np.abs(got - ref["logit"]).max() <= logit_tol(ref["logit"], FP32_REL)
    E, _ = m.gathered_block(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda())
    assert np.array_equal(E.cpu().numpy(), ref["E32"])
    # size-independent properties: batch-order equivariance and sample independence
    perm = np.random.default_rng(0).permutation(4096)
    got_p = run(m, Xi[perm], Xv[perm])
    assert np.array_equal(got_p, got[perm])
    assert np.array_equal(run(m, Xi[:100], Xv[:100]), got[:100])


def test_config3_pruned_full_size():
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = prune.one_shot_prune(synth.make_weights(cfg, seed=42), 0.9, 0.444, 1.0)
    Xi, Xv = synth.make_inputs(cfg, 4096, seed=3, dist="zipf")
    ref = closed_form.forward(cfg, w, Xi, Xv)["logit"]
    for precision in ("fp32", "fp32_csr"):
        got = run(to_cuda(cfg, w, precision=precision), Xi, Xv)
        assert np.abs(got - ref).max() <= logit_tol(ref, FP32_REL), precision


def test_config4_qr_full_cardinality_tables():
    """QR (mult, c=4, threshold 200) on the un-thresholded Kaggle cardinalities: 338 MB of tables."""
    cfg = PathConfig(39, synth.CRITEO_KAGGLE, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True,
                     qr_flag=1, qr_collisions=4, qr_threshold=200, deep_nodes=64)
    w = synth.make_weights(cfg, seed=11)
    Xi, Xv = synth.make_inputs(cfg, 2048, seed=4)
    ref = closed_form.forward(cfg, w, Xi, Xv)
    m = to_cuda(cfg, w)
    got = run(m, Xi, Xv)
    assert np.abs(got - ref["logit"]).max() <= logit_tol(ref["logit"], FP32_REL)
    E, _ = m.gathered_block(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda())
    assert np.array_equal(E.cpu().numpy(), ref["E32"])


def test_standalone_qr_lookup_bit_exact():
    from xsdeepfwfm_deprecated_b200.model import QREmbeddingBag
    torch.manual_seed(0)
    for op, c in (("mult", 4), ("add", 7)):
        t = QREmbeddingBag(100003, 10, c, operation=op, mode="sum").cuda()
        idx = torch.randint(0, 100003, (777, 1), device="cuda")
        got = t(idx)
        q, r = idx[:, 0] // c, idx[:, 0] % c
        want = t.weight_q[q] * t.weight_r[r] if op == "mult" else t.weight_q[q] + t.weight_r[r]
        assert torch.equal(got, want.detach())


@pytest.mark.parametrize("B", [1, 127, 128, 129, 1000, 20000])
def test_bf16_ragged_and_multi_tile(B):
    """Tensor-core path: partial 128-row tiles and more tiles than SMs (persistent loop, barrier phase wrap)."""
    c = load_case("deepfwfm_fwlw")
    cfg = c["cfg"]
    Xi, Xv = synth.make_inputs(cfg, B, seed=B)
    got = run(to_cuda(cfg, c["weights"], precision="bf16"), Xi, Xv)
    ref = closed_form.forward(cfg, c["weights"], Xi, Xv)
    assert got.shape == (B,)
    assert np.abs(got - ref["logit"]).max() <= logit_tol(ref["logit"], FP32_REL) + BF16_DEEP_REL * np.abs(ref["deep"]).max()


def test_bf16_equals_fp32_on_bf16_representable_problem():
    """With weights and inputs that are exactly representable in bf16 and a one-layer MLP the tensor path has no
    operand rounding at all: it must agree with the fp32 path to accumulation order."""
    cfg = PathConfig(39, [1] * 13 + [50] * 26, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True,
                     deep_nodes=400, h_depth=1)
    w = synth.make_weights(cfg, seed=3)
    for k in w:
        if "2nd_embeddings" in k or "net_1_linear" in k:
            w[k] = torch.from_numpy(w[k]).bfloat16().float().numpy()
    Xi, Xv = synth.make_inputs(cfg, 300, seed=9)
    Xv = np.minimum(Xv, 3.0).astype(np.float32)           # small integers keep W[0]*Xv exact in bf16
    for f in range(13):
        w[f"fm_2nd_embeddings.{f}.weight"] = (np.round(w[f"fm_2nd_embeddings.{f}.weight"] * 8) / 8).astype(np.float32)
    ref = closed_form.forward(cfg, w, Xi, Xv)["logit"]
    a = run(to_cuda(cfg, w, precision="bf16"), Xi, Xv)
    assert np.abs(a - ref).max() <= logit_tol(ref, FP32_REL)


# ----------------------------------------------------------------------------------------------------------------
# bf16x3: the fused single-kernel forward on split bf16 operands (csrc/fused_tc.cu).  It claims the fp32 bound.
@pytest.mark.parametrize("name", [n for n in CASES if load_case(n)["cfg"].use_deep])
def test_golden_logits_bf16x3(name):
    c = load_case(name)
    m = to_cuda(c["cfg"], c["weights"], precision="bf16x3")
    got = run(m, c["Xi"], c["Xv"])
    assert got.shape == c["logits"].shape and got.dtype == np.float32
    assert np.abs(got - c["logits"]).max() <= logit_tol(c["logits"], FP32_REL)
    ref64 = closed_form.forward(c["cfg"], c["weights"], c["Xi"], c["Xv"])["logit"]
    assert np.abs(got - ref64).max() <= logit_tol(ref64, FP32_REL)


def test_fused_kernel_takes_the_baseline_shapes():
    """The Criteo and Twitter shapes of BASELINE.json must run as the single fused kernel, not the staged path."""
    from xsdeepfwfm_deprecated_b200 import _lib
    lib = _lib.load()
    for name in ("deepfwfm_fwlw", "twitter_shape", "qr_mult_fwlw", "pruned", "k7", "deepfwfm_h4"):
        c = load_case(name)
        for precision in ("bf16", "bf16x3"):
            m = to_cuda(c["cfg"], c["weights"], precision=precision)
            plan = m._get_plan()
            plan.ensure_image(m, precision)
            assert lib.dfw_fused_supported(plan.model_ref, _lib.PRECISIONS[precision]) == 1, (name, precision)
            l0 = lib.dfw_launch_count()
            run(m, c["Xi"], c["Xv"])
            assert lib.dfw_launch_count() - l0 == 1, (name, precision)


@pytest.mark.parametrize("B", [1, 2, 31, 32, 33, 127, 129, 1000, 4097, 20000])
def test_bf16x3_ragged_and_multi_tile(B):
    """Partial 32-sample tiles, partial clusters, and more tiles than SMs (persistent loop, barrier phase wrap)."""
    c = load_case("deepfwfm_fwlw")
    cfg = c["cfg"]
    Xi, Xv = synth.make_inputs(cfg, B, seed=B)
    ref = closed_form.forward(cfg, c["weights"], Xi, Xv)
    for precision, rel, extra in (("bf16x3", FP32_REL, 0.0), ("bf16", FP32_REL, BF16_DEEP_REL * np.abs(ref["deep"]).max())):
        got = run(to_cuda(cfg, c["weights"], precision=precision), Xi, Xv)
        assert got.shape == (B,)
        err = np.abs(got - ref["logit"]).max()
        assert err <= logit_tol(ref["logit"], rel) + extra, (precision, B, float(err), int(np.abs(got - ref["logit"]).argmax()))


def test_config2_full_size_batch_4096_bf16x3():
    """BASELINE config 2 at full size on the fused kernel: fp32 bound, batch-order equivariance, sample independence."""
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = synth.make_weights(cfg, seed=42)
    Xi, Xv = synth.make_inputs(cfg, 4096, seed=0)
    ref = closed_form.forward(cfg, w, Xi, Xv)
    m = to_cuda(cfg, w, precision="bf16x3")
    got = run(m, Xi, Xv)
    assert np.abs(got - ref["logit"]).max() <= logit_tol(ref["logit"], FP32_REL)
    # the deep term alone (what the split operands touch), against the fp64 oracle
    shallow = ref["first"] + ref["second"] + float(w["bias"][0])
    assert np.abs((got - shallow) - ref["deep"]).max() <= 5e-5 * np.abs(ref["deep"]).max() + 1e-5 * np.abs(ref["logit"]).max()
    perm = np.random.default_rng(0).permutation(4096)
    assert np.array_equal(run(m, Xi[perm], Xv[perm]), got[perm])
    assert np.array_equal(run(m, Xi[:100], Xv[:100]), got[:100])
    with torch.no_grad():
        logits, prob = m(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda(), return_prob=True)
    assert torch.allclose(prob, torch.sigmoid(logits), atol=2e-7, rtol=1e-6)


def test_config3_pruned_and_config5_twitter_bf16x3():
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = prune.one_shot_prune(synth.make_weights(cfg, seed=42), 0.9, 0.444, 1.0)
    Xi, Xv = synth.make_inputs(cfg, 4096, seed=3, dist="zipf")
    ref = closed_form.forward(cfg, w, Xi, Xv)["logit"]
    got = run(to_cuda(cfg, w, precision="bf16x3"), Xi, Xv)
    assert np.abs(got - ref).max() <= logit_tol(ref, FP32_REL)
    c = load_case("twitter_shape")
    Xi, Xv = synth.make_inputs(c["cfg"], 3000, seed=8, xv="unit")
    ref = closed_form.forward(c["cfg"], c["weights"], Xi, Xv)["logit"]
    got = run(to_cuda(c["cfg"], c["weights"], precision="bf16x3"), Xi, Xv)
    assert np.abs(got - ref).max() <= logit_tol(ref, FP32_REL)


def test_streamed_host_inference_matches_forward():
    """dfw_forward_host_stream (3 rotating streams, ragged last batch) == forward + sigmoid, for every precision."""
    c = load_case("deepfwfm_fwlw")
    cfg = c["cfg"]
    Xi, Xv = synth.make_inputs(cfg, 5000, seed=21)
    for precision in ("fp32", "bf16x3", "bf16"):
        m = to_cuda(cfg, c["weights"], precision=precision)
        with torch.no_grad():
            logits, prob = m(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda(), return_prob=True)
        got_p, got_l = m.predict_proba_host(Xi, Xv, batch_size=512, want_logits=True, batches_in_flight=4)
        assert np.array_equal(got_l, logits.cpu().numpy()), precision
        assert np.array_equal(got_p, prob.cpu().numpy()), precision
        got_p2 = m.predict_proba_host(Xi, Xv, batch_size=8192)
        assert np.array_equal(got_p2, got_p), precision


@pytest.mark.parametrize("nodes,depth", [(300, 3), (136, 2), (64, 1), (16, 3), (250, 4), (512, 2), (520, 2)])
def test_fused_kernels_odd_widths_and_depths(nodes, depth):
    """Neuron-tile edge cases of the fused kernels: an odd number of 128-neuron tiles (the pair kernel's second CTA then has no
    tile in the last pair), widths that are not multiples of 16 or 64, one-tile layers (ring 1 idle), depth 1..4, and a width
    above the fused limit (520 > 512: staged fallback).  Criteo shape, so the CTA-pair kernel runs for B > 32."""
    cfg = PathConfig(39, [1] * 13 + [50 + 37 * i for i in range(26)], use_fm=False, use_fwfm=True, use_deep=True,
                     use_fwlw=True, deep_nodes=nodes, h_depth=depth)
    w = synth.make_weights(cfg, seed=nodes + depth)
    for B in (20, 333):
        Xi, Xv = synth.make_inputs(cfg, B, seed=B + nodes)
        ref = closed_form.forward(cfg, w, Xi, Xv)
        for precision, rel, extra in (("bf16x3", FP32_REL, 0.0), ("bf16", FP32_REL, BF16_DEEP_REL * np.abs(ref["deep"]).max())):
            got = run(to_cuda(cfg, w, precision=precision), Xi, Xv)
            assert np.abs(got - ref["logit"]).max() <= logit_tol(ref["logit"], rel) + extra, (precision, B)


def test_int32_index_format_is_bit_identical():
    """DFW_XI_INT32: the packed int32 input format gives the same bits as the reference's int64, through
    forward (fused and staged kernels) and through the streamed host path; a wrong index dtype is refused."""
    c = load_case("qr_mult_fwlw")
    cfg = c["cfg"]
    Xi, Xv = synth.make_inputs(cfg, 700, seed=5)
    for precision in ("bf16x3", "fp32"):
        m64 = to_cuda(cfg, c["weights"], precision=precision)
        m32 = to_cuda(cfg, c["weights"], precision=precision, index_dtype="int32")
        a = run(m64, Xi, Xv)
        with torch.no_grad():
            b = m32(torch.from_numpy(Xi.astype(np.int32)).cuda(), torch.from_numpy(Xv).cuda()).cpu().numpy()
        assert np.array_equal(a, b), precision
        pa = m64.predict_proba_host(Xi, Xv, batch_size=256)
        pb = m32.predict_proba_host(Xi.astype(np.int32), Xv, batch_size=256)
        assert np.array_equal(pa, pb), precision
        with pytest.raises(TypeError):
            m32(torch.from_numpy(Xi).cuda(), torch.from_numpy(Xv).cuda())


def test_concurrent_forwards_on_several_streams():
    """bench.py spreads independent forwards over 3 streams (and a server would): the same module object must give the same
    bits when its forwards overlap on the device."""
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = synth.make_weights(cfg, seed=42)
    m = to_cuda(cfg, w, precision="bf16x3")
    batches = [synth.make_inputs(cfg, 4096, seed=100 + i) for i in range(6)]
    dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in batches]
    with torch.no_grad():
        want = [m(a, b).clone() for a, b in dev]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(3)]
        got = [None] * len(dev)
        for rep in range(5):
            for i, (a, b) in enumerate(dev):
                with torch.cuda.stream(streams[i % 3]):
                    got[i] = m(a, b)
        torch.cuda.synchronize()
    for g, wnt in zip(got, want):
        assert torch.equal(g, wnt)


def test_host_transports_mapped_and_staged_are_bit_identical():
    """dfw_forward_host_stream picks its transport from the buffers: pinned host memory -> the fused kernel loads Xi / Xv and
    stores its results over PCIe itself ("mapped", one launch per batch); pageable memory -> staged cudaMemcpyAsync.  Same
    bits either way, int64 and int32 indices, ragged last batch, logits and probabilities."""
    import ctypes
    from xsdeepfwfm_deprecated_b200 import _lib
    lib = _lib.load()
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w = synth.make_weights(cfg, seed=42)
    N, bs = 4096 * 3 + 77, 4096
    Xi, Xv = synth.make_inputs(cfg, N, seed=31)
    for idt in ("int64", "int32"):
        m = to_cuda(cfg, w, precision="bf16x3", index_dtype=idt)
        plan = m._get_plan()
        plan.ensure_image(m, "bf16x3")
        prec = _lib.PRECISIONS["bf16x3"]
        ws = torch.zeros(lib.dfw_forward_host_stream_workspace_bytes(plan.model_ref, bs, prec) + 4096, dtype=torch.uint8,
                         device="cuda")
        xi_page = torch.from_numpy(np.ascontiguousarray(Xi[:, :, 0].astype(idt)))
        xv_page = torch.from_numpy(Xv.copy())
        outs = {}
        for name, pin in (("staged", False), ("mapped", True)):
            xi_h = xi_page.pin_memory() if pin else xi_page
            xv_h = xv_page.pin_memory() if pin else xv_page
            lo = torch.full((N,), -7.0)
            pr = torch.full((N,), -7.0)
            if pin:
                lo, pr = lo.pin_memory(), pr.pin_memory()
            args = (xi_h.data_ptr(), xv_h.data_ptr(), lo.data_ptr(), pr.data_ptr())
            assert lib.dfw_host_transport_is_mapped(plan.model_ref, prec, *args) == int(pin)
            _lib.check(lib.dfw_forward_host_stream(plan.model_ref, args[0], args[1], N, bs, prec, ws.data_ptr(), ws.numel(),
                                                   args[2], args[3], torch.cuda.current_stream().cuda_stream), name)
            outs[name] = (lo.clone(), pr.clone())
        assert torch.equal(outs["staged"][0], outs["mapped"][0]) and torch.equal(outs["staged"][1], outs["mapped"][1]), idt
        with torch.no_grad():
            idx = torch.from_numpy(Xi.astype(idt)).cuda()
            want = m(idx, torch.from_numpy(Xv).cuda()).cpu()
        assert torch.equal(outs["mapped"][0], want), idt
        # the fp32 (staged-kernel) precision cannot use the mapped transport and says so
        assert lib.dfw_host_transport_is_mapped(plan.model_ref, _lib.PRECISIONS["fp32"], *args) == 0


# ------------------------------------------------------------------------------- one-shot pruning on the device (8(f) row 3)
def _prune_case(cfg, seed, sparse, emb_r, emb_corr, precision="bf16x3"):
    w = synth.make_weights(cfg, seed=seed)
    want = prune.one_shot_prune(w, sparse, emb_r, emb_corr)
    m = to_cuda(cfg, w, precision=precision)
    report = m.prune_one_shot(sparse=sparse, emb_r=emb_r, emb_corr=emb_corr)
    got = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    return w, want, m, report, got


def test_device_pruning_is_bit_identical_to_the_reference_recipe_config3():
    """BASELINE config 3 at full size (1.33 M rows): thresholds equal to the reference bisection's doubles, every parameter
    bit-identical to the reference recipe's output, the paper's census, and the pruned model's logits inside the fp32 bound."""
    cfg = PathConfig(39, synth.CRITEO_PAPER, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=True)
    w, want, m, report, got = _prune_case(cfg, 42, 0.9, 0.444, 1.0)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    # thresholds: the device bisection takes the same fp64 steps as binary_search_threshold
    stacked = np.concatenate([w[k] for k in w if "fm_2nd_embeddings" in k], axis=0)
    assert report["emb"][0] == prune.bisect_threshold(stacked, 0.9 * 0.444, stacked.size)
    for k in ("net_1_linear_1.weight", "net_1_linear_2.weight", "net_1_linear_3.weight", "fwfm_linear.weight"):
        assert report[k][0] == prune.bisect_threshold(w[k], 0.9, w[k].size), k
        assert report[k][2] == int((want[k] == 0).sum())
    R = w["field_cov.weight"]
    assert report["field_cov.weight"][0] == prune.bisect_threshold(np.float32(0.5) * (R + R.T), 0.9, R.size)
    nz = sum(int((v != 0).sum()) for v in got.values())
    assert abs(nz - 8_012_094) / 8_012_094 < 1e-3          # the paper's D-90 % & R-90 % & F-40 % parameter count
    Xi, Xv = synth.make_inputs(cfg, 4096, seed=9, dist="zipf")
    ref = closed_form.forward(cfg, want, Xi, Xv)["logit"]
    for precision in ("bf16x3", "fp32_csr"):
        m.precision = precision
        assert np.abs(run(m, Xi, Xv) - ref).max() <= logit_tol(ref, FP32_REL), precision


@pytest.mark.parametrize("name,sparse,emb_r,emb_corr", [("deepfwfm_fwlw", 0.5, 0.6, 0.7), ("qr_mult_fwlw", 0.9, 0.444, 1.0),
                                                        ("deepfm", 0.3, 1.0, 1.0), ("twitter_shape", 0.95, 0.1, 0.5)])
def test_device_pruning_matches_the_recipe_on_the_golden_models(name, sparse, emb_r, emb_corr):
    """Other switch sets: QR tables (quotient and remainder tensors join the stacked embedding set), no fwfm_linear / no
    field_cov (DeepFM), the Twitter shape; a CUDA tensor handed to binary_search_threshold takes the device path."""
    c = load_case(name)
    w = c["weights"]
    want = prune.one_shot_prune(w, sparse, emb_r, emb_corr)
    m = to_cuda(c["cfg"], w)
    k0 = "net_1_linear_1.weight"
    t_dev = m.binary_search_threshold(m.net_1_linear_1.weight.data, sparse, w[k0].size)
    assert t_dev == prune.bisect_threshold(w[k0], sparse, w[k0].size)
    m.prune_one_shot(sparse=sparse, emb_r=emb_r, emb_corr=emb_corr)
    got = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    ref = closed_form.forward(c["cfg"], want, c["Xi"], c["Xv"])["logit"]
    assert np.abs(run(m, c["Xi"], c["Xv"]) - ref).max() <= logit_tol(ref, FP32_REL)


def test_device_pruning_edge_cases():
    """Rates the bisection cannot meet (all-equal magnitudes: 101 probes, like the reference), an all-zero tensor, a tensor
    whose storage is not 16-byte aligned, and the switches that leave parts untouched."""
    m = to_cuda(*(lambda c: (c["cfg"], c["weights"]))(load_case("deepfwfm_fwlw")))
    dev = m.bias.device
    for t in (torch.full((1000,), 0.25, device=dev), torch.zeros(777, device=dev),
              torch.randn(4099, device=dev)[3:], torch.randn(5, device=dev)):
        a = t.detach().cpu().numpy()
        for rate in (0.0, 0.5, 1.0):
            assert m.binary_search_threshold(t, rate, a.size) == prune.bisect_threshold(a, rate, a.size), (a.size, rate)
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    rep = m.prune_one_shot(sparse=0.8, prune_fm=0, prune_r=0, prune_deep=1)
    assert "emb" not in rep and "field_cov.weight" not in rep and "net_1_linear_2.weight" in rep
    after = m.state_dict()
    for k in before:
        changed = not torch.equal(before[k], after[k])
        assert changed == ("linear" in k and "weight" in k), k


def test_device_pruning_reproduces_the_reference_pruned_fixture():
    """tests/golden/pruned.npz was made by the REFERENCE: its own binary_search_threshold and masking on its own module, then
    its forward.  Pruning the same unpruned weights on the device must give the same weights (checksum of every parameter),
    the same non-zero census and, through the fused kernel, the reference's logits."""
    import json
    import os
    from golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "pruned.npz"))
    cfg = PathConfig.from_json(json.loads(str(z["cfg"])))
    w = synth.make_weights(cfg, seed=int(z["seed"]), emb_scale=float(z["emb_scale"]))        # unpruned
    assert synth.weights_checksum(w) != str(z["checksum"])
    for precision, rel in (("fp32", FP32_REL), ("bf16x3", FP32_REL), ("fp32_csr", FP32_REL)):
        m = to_cuda(cfg, w, precision=precision)
        m.prune_one_shot(sparse=0.9, emb_r=0.444, emb_corr=1.0)
        got_w = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        assert synth.weights_checksum(got_w) == str(z["checksum"]), precision
        for k, n in json.loads(str(z["nnz"])).items():
            assert int(np.count_nonzero(got_w[k])) == n, k
        got = run(m, z["Xi"].astype(np.int64), z["Xv"].astype(np.float32))
        assert np.abs(got - z["logits"]).max() <= logit_tol(z["logits"], rel), precision


def test_cached_loader_columns_feed_the_host_path(tmp_path):
    """the binary cache of utils.data_preprocess (int32 indices, memory-mapped) goes through
    predict_proba_host / eval_by_batch unchanged and gives the same bits as in-memory int64 arrays, for both index formats."""
    from xsdeepfwfm_deprecated_b200.utils import data_preprocess as dp
    t = load_tiny()
    v = t["variants"]["deep_fwlw"]
    w = tiny_weights(v)
    n = 3000
    Xi, Xv, y = t["Xi"][:n], t["Xv"][:n], t["y"][:n]
    path = dp.write_cache(dict(index=Xi[:, :, 0], value=Xv, label=y.astype(np.int8), feature_sizes=v["cfg"].feature_sizes),
                          str(tmp_path / "cache"))
    c = dp.read_cache(path)
    assert isinstance(c["index"], np.memmap) and c["index"].dtype == np.int32
    for idt in ("int64", "int32"):
        m = to_cuda(v["cfg"], w, precision="bf16x3", index_dtype=idt)
        want = m.predict_proba_host(Xi, Xv, batch_size=1024)
        got = m.predict_proba_host(c["index"], c["value"], batch_size=1024)
        assert np.array_equal(got, want), idt
        a = m.eval_by_batch(Xi, Xv, y, n)
        b = m.eval_by_batch(c["index"], c["value"], c["label"], n)
        assert a == b, idt
