"""Drop-in ``DeepFMs`` module whose ``forward(Xi, Xv)`` runs on hand-written sm_100a kernels.

Mirrors the inference-facing API of the reference class (model/DeepFMs.py:47-1102):

* constructor keywords and defaults ............ model/DeepFMs.py:81-89
* parameter names / shapes (``state_dict``) ..... model/DeepFMs.py:185-222, 246-283, 1066-1091
* ``forward(Xi, Xv) -> logits (B,)`` ............ model/DeepFMs.py:285-469
* ``init_weights`` ............................... model/DeepFMs.py:472-495
* ``eval_by_batch`` / ``predict*`` / ``evaluate`` . model/DeepFMs.py:750-784, 848-903
* ``binary_search_threshold`` ................... model/DeepFMs.py:807-823

so a reference checkpoint (dense, or magnitude-pruned = zeros inside dense tensors) loads with
``load_state_dict`` unchanged and ``utils.util.get_model`` can construct it.  Differences, on
purpose:

* no CPU fallback -- without an sm_100 GPU ``forward`` raises (the reference silently falls
  back to CPU, model/DeepFMs.py:153-155);
* unsupported switch combinations raise ``ValueError`` instead of ``exit(1)``
  (model/DeepFMs.py:159-161, 178-180): FFM, ``use_logit``, ``num_deeps != 1``, batch norm, the
  quantisation switches, ``qr_operation='concat'``, the deep-only model;
* inference only: ``forward`` in ``train()`` mode with dropout, ``fit`` and the KD helpers raise;
* ``logger=None`` is accepted.

The kernels read the embedding tables and the fp32 MLP weights in place (pointer table in device
memory), so in-place edits of those are seen by the next call.  Derived images -- the shallow image
(symmetrised ``field_cov`` / pair list, ``fwfm_linear`` x ``fm_1st``, descriptors), bf16 MLP weights
for the tcgen05 path, CSR for the pruned-MLP path -- are rebuilt whenever the module is moved,
re-loaded, re-initialised, switched with ``train()/eval()`` (every reference inference entry point
calls ``eval()`` first, model/DeepFMs.py:757, 859, 870) or on ``repack()``.
"""
from __future__ import annotations

import ctypes as C
import logging
import random
from typing import Optional

import numpy as np
import torch
from torch import nn
from sklearn.metrics import roc_auc_score, log_loss, precision_recall_curve, auc as _sk_auc

from .. import _lib
from .QREmbeddingBag import QREmbeddingBag

_NULL_LOGGER = logging.getLogger("xsdeepfwfm_deprecated_b200.null")
_NULL_LOGGER.addHandler(logging.NullHandler())
_NULL_LOGGER.propagate = False


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _Plan:
    """Pointer table + derived images for one (device, parameter-storage) state of a module."""

    def __init__(self, owner: "DeepFMs"):
        lib = _lib.load()
        dev = owner.bias.device
        if dev.type != "cuda":
            raise RuntimeError(
                "DeepFMs.forward needs the module on an sm_100 CUDA device; this implementation has no CPU "
                "fallback (the reference would silently run on CPU here, model/DeepFMs.py:153-155)")
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        self.device = dev
        self.keep = []          # tensors that must outlive the plan
        F, K, num = owner.field_size, owner.embedding_size, owner.num
        params = []

        def chk(p: torch.Tensor, name: str) -> torch.Tensor:
            t = p.data
            if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError(f"parameter {name} must be contiguous fp32 on {dev} (got {t.dtype}, {t.device})")
            params.append(p)        # the Parameter itself: re-binding .data must be noticed by stale()
            return t

        descs = (_lib.FieldDesc * F)()
        for f in range(F):
            d = descs[f]
            n = int(owner.feature_sizes[f])
            if n >= 2 ** 31:
                raise ValueError("tables with >= 2^31 rows are not supported")
            d.rows, d.collisions, d.n_ranks = n, 1, 0
            e2 = owner.fm_2nd_embeddings[f]
            if isinstance(e2, QREmbeddingBag):
                d.w2 = chk(e2.weight_q, f"fm_2nd_embeddings.{f}.weight_q").data_ptr()
                d.w2_r = chk(e2.weight_r, f"fm_2nd_embeddings.{f}.weight_r").data_ptr()
                d.collisions = e2.num_collisions
                d.qr_op = _lib.TABLE_QR_MULT if e2.operation == "mult" else _lib.TABLE_QR_ADD
            else:
                d.w2 = chk(e2.weight, f"fm_2nd_embeddings.{f}.weight").data_ptr()
                d.qr_op = _lib.TABLE_PLAIN
            if not owner.use_fwlw:
                e1 = owner.fm_1st_embeddings[f]
                if isinstance(e1, QREmbeddingBag):
                    d.w1 = chk(e1.weight_q, f"fm_1st_embeddings.{f}.weight_q").data_ptr()
                    d.w1_r = chk(e1.weight_r, f"fm_1st_embeddings.{f}.weight_r").data_ptr()
                    d.collisions = e1.num_collisions
                    d.qr1_op = _lib.TABLE_QR_MULT if e1.operation == "mult" else _lib.TABLE_QR_ADD
                else:
                    d.w1 = chk(e1.weight, f"fm_1st_embeddings.{f}.weight").data_ptr()
                    d.qr1_op = _lib.TABLE_PLAIN
        owner._patch_field_descs(descs, self)
        host = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8)
        self.fields_dev = host.to(dev)
        self.keep.append(self.fields_dev)

        m = _lib.Model()
        m.struct_bytes = C.sizeof(_lib.Model)
        m.abi_version = _lib.DFW_ABI_VERSION
        flags = 0
        if owner.use_fwfm:
            flags |= _lib.USE_FWFM
            m.field_cov = chk(owner.field_cov.weight, "field_cov.weight").data_ptr()
        if owner.use_fwlw:
            flags |= _lib.USE_FWLW
            m.fwfm_linear = chk(owner.fwfm_linear.weight, "fwfm_linear.weight").data_ptr()
        if owner.use_lw:
            flags |= _lib.USE_LW
            m.fm_1st = chk(owner.fm_1st.weight, "fm_1st.weight").data_ptr()
        if owner.check_index:
            flags |= _lib.CHECK_INDEX
        if getattr(owner, "index_dtype", "int64") == "int32":
            flags |= _lib.XI_INT32
        if getattr(owner, "throughput_hint", False):
            flags |= _lib.HINT_THROUGHPUT
        m.field_size, m.numerical, m.embedding_size = F, num, K
        m.fields = self.fields_dev.data_ptr()
        m.bias = chk(owner.bias, "bias").data_ptr()
        if owner.use_deep:
            flags |= _lib.USE_DEEP
            m.depth = owner.h_depth
            in_dim = F * K
            for l in range(owner.h_depth):
                lin = getattr(owner, f"net_1_linear_{l + 1}")
                m.widths[l] = lin.out_features
                m.W[l] = chk(lin.weight, f"net_1_linear_{l + 1}.weight").data_ptr()
                m.b[l] = chk(lin.bias, f"net_1_linear_{l + 1}.bias").data_ptr()
                assert lin.in_features == in_dim
                in_dim = lin.out_features
            m.fc = chk(owner.net_1_fc.weight, "net_1_fc.weight").data_ptr()
        m.flags = flags
        self.model = m
        self.model_ref = C.byref(m)
        self.params = params
        self.ptrs = tuple((p.data_ptr(), p._version) for p in params)
        # shallow image: U / pair list / fwlw weights / descriptors in the kernel's shared-memory layout
        nimg = lib.dfw_shallow_image_bytes(self.model_ref)
        self.shallow_image = torch.zeros(nimg, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.dfw_pack_shallow(self.model_ref, self.shallow_image.data_ptr(), _stream_ptr(dev)),
                       "dfw_pack_shallow")
        m.shallow_image = self.shallow_image.data_ptr()
        if owner.use_fwfm:      # host snapshot of field_cov: the fused kernel takes the field matrix as a kernel parameter
            self.field_cov_host = owner.field_cov.weight.detach().to("cpu", torch.float32).contiguous()
            m.field_cov_host = self.field_cov_host.data_ptr()
        self.images = set()     # which derived images exist: "bf16", "csr"
        self.workspace = None
        self.host_ws = None

    def stale(self) -> bool:
        """Re-bound storage (init_weights, .cuda()) or an in-place edit that autograd's version counter sees (copy_, mul_, an
        optimizer step).  Edits through ``param.data[...] = ...`` (the reference's pruner) bypass that counter: tables and fp32
        MLP weights are read in place anyway, the derived images (shallow image, bf16 / CSR weights) need a train() -> eval()
        transition (what the reference's fit() / eval_by_batch() flow does) or repack()."""
        return self.ptrs != tuple((p.data_ptr(), p._version) for p in self.params)

    # -- derived images ------------------------------------------------------------------
    def ensure_image(self, owner: "DeepFMs", precision: str):
        if not owner.use_deep or precision == "fp32" or precision in self.images:
            return
        lib = _lib.load()
        st = _stream_ptr(self.device)
        m = self.model
        in_dim = owner.field_size * owner.embedding_size
        for l in range(owner.h_depth):
            lin = getattr(owner, f"net_1_linear_{l + 1}")
            out_dim = lin.out_features
            if precision in ("bf16", "bf16x3"):
                # hi = bf16(W); bf16x3 adds lo = bf16(W - hi) (split operands, see csrc/fused_tc.cu)
                nbytes = lib.dfw_pack_mlp_bf16_bytes(out_dim, in_dim)
                hi = lo = None
                if not m.Wbf16[l]:
                    hi = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                if precision == "bf16x3":
                    lo = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                if hi is not None or lo is not None:
                    _lib.check(lib.dfw_pack_mlp_bf16_split(lin.weight.data_ptr(), out_dim, in_dim, _ptr(hi), _ptr(lo), st),
                               "dfw_pack_mlp_bf16_split")
                if hi is not None:
                    m.Wbf16[l] = hi.data_ptr()
                    self.keep.append(hi)
                if lo is not None:
                    m.Wbf16_lo[l] = lo.data_ptr()
                    self.keep.append(lo)
            else:  # fp32_csr
                row_ptr = torch.empty(out_dim + 1, dtype=torch.int32, device=self.device)
                _lib.check(lib.dfw_csr_count(lin.weight.data_ptr(), out_dim, in_dim, row_ptr.data_ptr(), st),
                           "dfw_csr_count")
                nnz = int(row_ptr[-1].item())          # packing is rare; one sync here is fine
                col = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
                val = torch.empty(max(nnz, 1), dtype=torch.float32, device=self.device)
                _lib.check(lib.dfw_csr_fill(lin.weight.data_ptr(), out_dim, in_dim, row_ptr.data_ptr(),
                                            col.data_ptr(), val.data_ptr(), st), "dfw_csr_fill")
                c = m.csr[l]
                c.row_ptr, c.col, c.val, c.nnz = row_ptr.data_ptr(), col.data_ptr(), val.data_ptr(), nnz
                c.max_row_nnz = int((row_ptr[1:] - row_ptr[:-1]).max().item())
                self.keep += [row_ptr, col, val]
            in_dim = out_dim
        self.images.add(precision)

    def get_workspace(self, nbytes: int) -> torch.Tensor:
        if self.workspace is None or self.workspace.numel() < nbytes:
            self.workspace = torch.zeros(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=self.device)
        return self.workspace


class DeepFMs(nn.Module):
    def __init__(self, field_size, feature_sizes, embedding_size=10, is_shallow_dropout=True, dropout_shallow=[0.0, 0.0],
                 h_depth=3, deep_nodes=400, is_deep_dropout=True, dropout_deep=[0.5, 0.5, 0.5, 0.5],
                 eval_metric=roc_auc_score, n_epochs=64, batch_size=2048, learning_rate=0.001, momentum=0.9,
                 optimizer_type='adam', is_batch_norm=False, verbose=False, random_seed=0, weight_decay=0.0,
                 use_fm=True, use_fwlw=False, use_lw=False, use_ffm=False, use_fwfm=False, use_deep=True,
                 loss_type='logloss',
                 use_cuda=True, n_class=1, greater_is_better=True, sparse=0.9, warm=10, num_deeps=1, numerical=13,
                 use_logit=0, embedding_bag=False, quantization_aware=False, dynamic_quantization=False,
                 static_quantization=False, static_calibrate=False,
                 qr_flag=0, qr_operation="mult", qr_collisions=1, qr_threshold=200, md_flag=0, md_threshold=200,
                 logger=None, precision="fp32", check_index=False, index_dtype="int64", throughput_hint=False):
        super().__init__()
        self.field_size = field_size
        self.feature_sizes = feature_sizes
        self.embedding_size = embedding_size
        self.is_shallow_dropout = is_shallow_dropout
        self.dropout_shallow = dropout_shallow
        self.h_depth = h_depth
        self.num_deeps = num_deeps
        self.deep_layers = [deep_nodes] * h_depth
        self.is_deep_dropout = is_deep_dropout
        self.dropout_deep = [0.5] * (h_depth + 1)
        self.n_epochs = n_epochs
        self.batch_size = batch_size
        self.learning_rate = learning_rate
        self.momentum = momentum
        self.optimizer_type = optimizer_type
        self.is_batch_norm = is_batch_norm
        self.verbose = verbose
        self.weight_decay = weight_decay
        self.random_seed = random_seed
        self.use_fm = use_fm
        self.use_fwlw = use_fwlw
        self.use_lw = use_lw
        self.use_ffm = use_ffm
        self.use_fwfm = use_fwfm
        self.use_logit = use_logit
        self.use_deep = use_deep
        self.loss_type = loss_type
        self.eval_metric = eval_metric
        self.use_cuda = use_cuda
        self.n_class = n_class
        self.greater_is_better = greater_is_better
        self.target_sparse = sparse
        self.warm = warm
        self.num = numerical
        self.embedding_bag = embedding_bag if not qr_flag else qr_flag  # qr needs embedding bag (:125)
        self.quantization_aware = quantization_aware
        self.static_quantization = static_quantization
        self.static_calibrate = static_calibrate
        self.dynamic_quantization = dynamic_quantization
        self.qr_flag = qr_flag
        self.qr_operation = qr_operation
        self.qr_collisions = qr_collisions
        self.qr_threshold = qr_threshold
        self.md_flag = md_flag
        self.md_threshold = md_threshold
        self.logger = logger if logger is not None else _NULL_LOGGER
        # extensions (keyword-only in spirit): MLP arithmetic and debug bounds checking
        self.precision = precision
        self.check_index = check_index
        # "int64" is what the reference feeds (torch.LongTensor); "int32" is the packed input format: half the
        # index bytes over PCIe.  Every Xi given to forward / predict_proba_host must then be int32.
        if index_dtype not in ("int64", "int32"):
            raise ValueError("index_dtype must be 'int64' or 'int32'")
        self.index_dtype = index_dtype
        # True: several forwards of this module are kept in flight (serving loop, several streams): the fused kernel trades the
        # latency of a lone launch for SM time per batch (DFW_HINT_THROUGHPUT); results are bit-identical.  Call repack() after
        # changing it.
        self.throughput_hint = bool(throughput_hint)
        self._plan: Optional[_Plan] = None
        self._frozen = False

        np.random.seed(self.random_seed)
        random.seed(self.random_seed)
        torch.manual_seed(self.random_seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed(self.random_seed)

        # ---- what the B200 hot path does not cover: refuse loudly ----
        if int(bool(use_fm)) + int(bool(use_ffm)) + int(bool(use_fwfm)) + int(bool(use_logit)) > 1:
            raise ValueError("only support one type only, please make sure to choose only LR, FM, FFM or FwFM part")
        if use_ffm:
            raise ValueError("use_ffm (field-aware FM) is outside the B200 hot path")
        if use_logit:
            raise ValueError("use_logit (plain logistic regression) is outside the B200 hot path")
        if not (use_fm or use_fwfm):
            if use_deep:
                raise ValueError("the deep-only model (model/DeepFMs.py:401-403) is outside the B200 hot path")
            raise ValueError("You have to choose more than one of (fm, ffm, fwfm, deep) models to use")
        if num_deeps != 1:
            raise ValueError("num_deeps != 1 is not supported")
        if is_batch_norm:
            raise ValueError("is_batch_norm is not supported (not reachable from utils.util.get_model either)")
        if quantization_aware or dynamic_quantization or static_quantization or static_calibrate:
            raise ValueError("the torch.quantization switches are CPU int8 paths and are not supported")
        if md_flag:
            raise ValueError("md_flag is not supported")
        if qr_flag and qr_operation not in ("mult", "add"):
            raise ValueError("qr_operation 'concat' changes the embedding width and is not supported")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if len(feature_sizes) != field_size:
            raise ValueError("len(feature_sizes) != field_size")
        if self.use_cuda and not torch.cuda.is_available():
            raise RuntimeError("use_cuda=True but no CUDA device is available; there is no CPU fallback "
                               "(construct with use_cuda=False only to inspect or load parameters)")
        self.logger.info("The model is %s%s", "deep" if use_deep else "", "fwfm" if use_fwfm else "fm")

        # ---- parameters, registered in the reference's order (model/DeepFMs.py:185-283) -----------
        self.bias = nn.Parameter(torch.Tensor([0.01]))
        if not self.use_fwlw:
            if self.embedding_bag:
                self.fm_1st_embeddings = self.create_emb(1, np.array(self.feature_sizes), sparse=False)
            else:
                self.fm_1st_embeddings = nn.ModuleList([nn.Embedding(n, 1) for n in self.feature_sizes])
        if self.dropout_shallow:
            self.fm_first_order_dropout = nn.Dropout(self.dropout_shallow[0])
        if self.embedding_bag:
            self.fm_2nd_embeddings = self.create_emb(self.embedding_size, np.array(self.feature_sizes), sparse=False)
        else:
            self.fm_2nd_embeddings = nn.ModuleList([nn.Embedding(n, self.embedding_size) for n in self.feature_sizes])
        if self.dropout_shallow:
            self.fm_second_order_dropout = nn.Dropout(self.dropout_shallow[1])
        if self.use_lw:
            self.fm_1st = nn.Linear(self.field_size, 1, bias=False)
        if self.use_fwlw:
            self.fwfm_linear = nn.Linear(self.embedding_size, self.field_size, bias=False)
        if self.use_fwfm:
            self.field_cov = nn.Linear(field_size, field_size, bias=False)
        if self.use_deep:
            if not self.use_fm:
                # the reference builds the second-order tables a second time for fwfm+deep
                # (model/DeepFMs.py:250-256); repeated so the RNG stream and initial values match
                if self.embedding_bag:
                    self.fm_2nd_embeddings = self.create_emb(self.embedding_size, np.array(self.feature_sizes),
                                                             sparse=False)
                else:
                    self.fm_2nd_embeddings = nn.ModuleList(
                        [nn.Embedding(n, self.embedding_size) for n in self.feature_sizes])
            if self.is_deep_dropout:
                self.net_1_linear_0_dropout = nn.Dropout(self.dropout_deep[0])
            widths = [self.field_size * self.embedding_size] + self.deep_layers
            for i in range(1, h_depth + 1):
                setattr(self, f"net_1_linear_{i}", nn.Linear(widths[i - 1], widths[i]))
                setattr(self, f"net_1_linear_{i}_relu", nn.ReLU())
                if self.is_deep_dropout:
                    setattr(self, f"net_1_linear_{i}_dropout", nn.Dropout(self.dropout_deep[i]))
            self.net_1_fc = nn.Linear(self.deep_layers[-1], 1, bias=False)

    # ------------------------------------------------------------------ construction helpers
    def create_emb(self, m, ln, sparse=True):
        """model/DeepFMs.py:1066-1091: QR table above the threshold, else EmbeddingBag with U(+-sqrt(1/n))."""
        emb_l = nn.ModuleList()
        for i in range(0, ln.size):
            n = int(ln[i])
            if self.qr_flag and n > self.qr_threshold:
                EE = QREmbeddingBag(n, m, self.qr_collisions, operation=self.qr_operation, mode="sum", sparse=sparse)
            else:
                EE = nn.EmbeddingBag(n, m, mode="sum", sparse=sparse)
                W = np.random.uniform(low=-np.sqrt(1 / n), high=np.sqrt(1 / n), size=(n, m)).astype(np.float32)
                EE.weight.data = torch.tensor(W, requires_grad=True)
            emb_l.append(EE)
        return emb_l

    def init_weights(self):
        """Same distributions and RNG call order as the reference (model/DeepFMs.py:472-495)."""
        model = self.train()
        require_update = True
        last_layer_size = 0
        on_cuda = self.bias.is_cuda

        def normal(size):
            t = torch.empty(size, dtype=torch.float32, device="cuda" if on_cuda else "cpu")
            return t.normal_()

        glorot = 1.0
        for name, param in model.named_parameters():
            if '1st_embeddings' in name:
                param.data = normal(param.data.size())
            elif '2nd_embeddings' in name:
                param.data = normal(param.data.size()).mul(0.01)
            elif 'linear' in name:
                if 'weight' in name:
                    glorot = np.sqrt(2.0 / np.sum(param.data.shape))
                param.data = normal(param.data.size()).mul(glorot)
            elif 'field_cov.weight' == name:
                param.data = normal(param.data.size()).mul(np.sqrt(2.0 / self.field_size / 2))
            else:
                if (self.use_fwfm or self.use_fm) and require_update:
                    last_layer_size += (self.field_size + self.embedding_size)
                if self.use_deep and require_update:
                    last_layer_size += (self.deep_layers[-1] + 1)
                require_update = False
                if name in ['fm_1st.weight', 'fm_2nd.weight'] or 'fc.weight' in name:
                    param.data = normal(param.data.size()).mul(np.sqrt(2.0 / last_layer_size))
        self.repack()

    # ------------------------------------------------------------------ plan / cache management
    def _patch_field_descs(self, descs, plan):
        """Hook for the rank-sharded subclass (sharded.py); single-GPU tables need nothing."""

    def repack(self):
        """Drop the pointer table and every derived weight image; rebuilt on the next forward."""
        self._plan = None
        return self

    def freeze(self, frozen: bool = True):
        """Skip the per-call parameter-pointer check (serving mode).  ``repack()`` un-stales manually."""
        self._frozen = frozen
        return self

    def _apply(self, fn, *a, **kw):
        self._plan = None
        self._host_ws = None
        return super()._apply(fn, *a, **kw)

    def train(self, mode: bool = True):
        # The plan (pointer table, packed weight images, workspace) is dropped only when the mode CHANGES: the reference's flow is
        # fit() [train mode: param.data[mask] = 0 pruning, invisible to the version counters] -> eval() -> inference, and that
        # transition must rebuild the images.  predict_proba / eval_by_batch call eval() on every call; rebuilding the plan there
        # cost milliseconds around a ~30 us kernel.  `.data[...]` edits made while staying in eval mode need repack().
        if bool(mode) != self.training:
            self._plan = None
        return super().train(mode)

    def load_state_dict(self, *a, **kw):
        self._plan = None
        return super().load_state_dict(*a, **kw)

    def cuda(self, device=None):
        self.use_cuda = True
        return super().cuda(device)

    def cpu(self):
        self.use_cuda = False
        return super().cpu()

    def _get_plan(self) -> _Plan:
        p = self._plan
        if p is None or (not self._frozen and p.stale()):
            p = self._plan = _Plan(self)
        return p

    # ------------------------------------------------------------------ the hot path
    def forward(self, Xi, Xv, return_prob: bool = False):
        """Xi (B, field_size - numerical, 1) int64, Xv (B, numerical) fp32, both on the module's device.

        Returns logits (B,) fp32 -- the reference's ``total_sum`` (model/DeepFMs.py:458-469).  With
        ``return_prob=True`` also returns sigmoid(logits), fused into the last kernel.

        Streams: the fused single-kernel path (``bf16x3`` / ``bf16`` on the dataset shapes) uses no workspace, so forwards of one
        module may overlap on several streams.  The staged paths (``fp32``, ``fp32_csr``, shapes outside the fused kernel)
        keep E and the activations in ONE per-module workspace: run those on one stream at a time.
        """
        if self.training and self.use_deep and self.is_deep_dropout:
            raise RuntimeError("this implementation is inference-only: call .eval() first (train-mode dropout "
                               "and backward are outside the B200 hot path)")
        plan = self._get_plan()
        dev = plan.device
        C_ = self.field_size - self.num
        if Xi.device != dev or Xv.device != dev:
            raise RuntimeError(f"inputs must be on {dev} (Xi on {Xi.device}, Xv on {Xv.device})")
        want = torch.int32 if self.index_dtype == "int32" else torch.int64
        if Xi.dtype != want:
            raise TypeError(f"Xi must be {want} (index_dtype={self.index_dtype!r}; the reference feeds torch.LongTensor)")
        if Xv.dtype != torch.float32:
            raise TypeError("Xv must be float32")
        B = Xi.shape[0]
        if Xi.dim() != 3 or Xi.shape[1] != C_ or Xi.shape[2] != 1:
            raise ValueError(f"Xi must have shape (B, {C_}, 1), got {tuple(Xi.shape)}")
        if Xv.dim() != 2 or Xv.shape[0] != B or Xv.shape[1] < self.num:
            raise ValueError(f"Xv must have shape (B, {self.num}), got {tuple(Xv.shape)}")
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        prob = torch.empty(B, dtype=torch.float32, device=dev) if return_prob else None
        if B == 0:
            return (logits, prob) if return_prob else logits
        precision = self.precision
        plan.ensure_image(self, precision)
        lib = _lib.load()
        prec = _lib.PRECISIONS[precision]
        nbytes = lib.dfw_forward_workspace_bytes(plan.model_ref, B, prec)
        ws = plan.get_workspace(nbytes)
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self.forward(Xi, Xv, return_prob)
        if self.check_index:
            ws[:4].zero_()
        rc = lib.dfw_forward(plan.model_ref, Xi.data_ptr(), Xi.stride(0), Xi.stride(1),
                             Xv.data_ptr() if self.num else None, Xv.stride(0), Xv.stride(1), B, prec,
                             ws.data_ptr(), ws.numel(), logits.data_ptr(), _ptr(prob), ws.data_ptr(),
                             _stream_ptr(dev))
        _lib.check(rc, "dfw_forward")
        if self.check_index:
            bad = int(ws[:4].view(torch.int32).item())
            if bad:
                raise IndexError(f"index out of range in field {bad - 1} (table has "
                                 f"{self.feature_sizes[bad - 1]} rows)")
        return (logits, prob) if return_prob else logits

    def gathered_block(self, Xi, Xv):
        """The (B, F, K) embedding block E the kernels build (== torch.stack(fm_2nd_emb_arr, 1),
        model/DeepFMs.py:337) plus the shallow partial sum -- exposed for the bit-exactness tests."""
        plan = self._get_plan()
        dev, B = plan.device, Xi.shape[0]
        FK = self.field_size * self.embedding_size
        E = torch.empty(B, FK, dtype=torch.float32, device=dev)
        shallow = torch.empty(B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().dfw_embed_fwfm(plan.model_ref, Xi.data_ptr(), Xi.stride(0), Xi.stride(1),
                                            Xv.data_ptr() if self.num else None, Xv.stride(0), Xv.stride(1), B,
                                            E.data_ptr(), FK, None, 0, shallow.data_ptr(), None, _stream_ptr(dev))
        _lib.check(rc, "dfw_embed_fwfm")
        return E.view(B, self.field_size, self.embedding_size), shallow

    # ------------------------------------------------------------------ host-buffer inference (e2e path)
    def predict_proba_host(self, Xi_np: np.ndarray, Xv_np: np.ndarray, batch_size: int = 8192,
                           want_logits: bool = False, batches_in_flight: int = 16):
        """numpy in, numpy out through ``dfw_forward_host_stream``: pinned staging, then per batch H2D, fused forward +
        sigmoid, D2H on rotating streams (copies overlap kernels), one synchronisation per ``batches_in_flight``
        batches -- what eval_by_batch / predict_proba do around forward (model/DeepFMs.py:771-777)."""
        plan = self._get_plan()
        dev = plan.device
        lib = _lib.load()
        prec = _lib.PRECISIONS[self.precision]
        plan.ensure_image(self, self.precision)
        n = len(Xi_np)
        C_ = self.field_size - self.num
        t_idx = torch.int32 if self.index_dtype == "int32" else torch.int64
        # no whole-array conversion: memory-mapped columns (utils.data_preprocess.read_cache) are sliced per chunk and cast
        # while they are copied into the pinned staging buffers
        Xi_np = np.asarray(Xi_np).reshape(n, C_)
        Xv_np = np.asarray(Xv_np).reshape(n, -1)
        bs = min(batch_size, max(n, 1))
        chunk = bs * max(1, min(batches_in_flight, -(-max(n, 1) // bs)))
        # the pinned staging buffers and the device workspace belong to the module, not to the plan: eval() (which every
        # inference entry point of the reference calls first) drops the plan, and re-pinning four buffers per call costs milliseconds
        hw = getattr(self, "_host_ws", None)
        if hw is None or hw[0] != (bs, prec, str(dev), self.index_dtype) or hw[2] < chunk:
            nbytes = lib.dfw_forward_host_stream_workspace_bytes(plan.model_ref, bs, prec)
            self._host_ws = ((bs, prec, str(dev), self.index_dtype), None, chunk,
                             torch.zeros(nbytes + 4096, dtype=torch.uint8, device=dev),
                            torch.empty(chunk * C_, dtype=t_idx).pin_memory(),
                            torch.empty(chunk * max(self.num, 1), dtype=torch.float32).pin_memory(),
                            torch.empty(chunk, dtype=torch.float32).pin_memory(),
                            torch.empty(chunk, dtype=torch.float32).pin_memory())
        _, _, chunk, ws, pxi, pxv, pprob, plogit = self._host_ws
        out = np.empty(n, dtype=np.float32)
        out_logit = np.empty(n, dtype=np.float32) if want_logits else None
        st = _stream_ptr(dev)
        with torch.cuda.device(dev):
            for o in range(0, n, chunk):
                e = min(n, o + chunk)
                b = e - o
                pxi.numpy()[:b * C_].reshape(b, C_)[...] = Xi_np[o:e]
                if self.num:
                    pxv.numpy()[:b * self.num].reshape(b, self.num)[...] = Xv_np[o:e, :self.num]
                rc = lib.dfw_forward_host_stream(plan.model_ref, pxi.data_ptr(), pxv.data_ptr(), b, bs, prec,
                                                 ws.data_ptr(), ws.numel(), plogit.data_ptr() if want_logits else None,
                                                 pprob.data_ptr(), st)
                _lib.check(rc, "dfw_forward_host_stream")
                out[o:e] = pprob.numpy()[:b]
                if want_logits:
                    out_logit[o:e] = plogit.numpy()[:b]
        return (out, out_logit) if want_logits else out

    # ------------------------------------------------------------------ the reference's inference harness
    def eval_by_batch(self, Xi, Xv, y, x_size):
        """(loss, metric, prauc, rce) like model/DeepFMs.py:750-784: slices of 8192, sigmoid, BCE, sklearn."""
        self.eval()
        Xi = np.asarray(Xi)[:x_size]
        Xv = np.asarray(Xv, dtype=np.float32)[:x_size]
        y = np.asarray(y, dtype=np.float64)[:x_size]
        prob, logit = self.predict_proba_host(Xi, Xv, batch_size=8192, want_logits=True)
        z = logit.astype(np.float64)
        # binary_cross_entropy_with_logits, mean over the set (the reference sums per-batch means * size)
        loss = float(np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))
        y_pred = prob.astype("float64")
        return loss, self.eval_metric(y, y_pred), self.compute_prauc(y_pred, y), self.compute_rce(y_pred, y)

    def compute_prauc(self, pred, gt):
        prec, recall, _ = precision_recall_curve(gt, pred)
        return _sk_auc(recall, prec)

    def calculate_ctr(self, gt):
        gt = np.asarray(gt)
        return float((gt == 1).sum()) / float(len(gt))

    def compute_rce(self, pred, gt):
        cross_entropy = log_loss(gt, pred)
        data_ctr = self.calculate_ctr(gt)
        strawman = log_loss(gt, [data_ctr for _ in range(len(gt))])
        return (1.0 - cross_entropy / strawman) * 100.0

    def binary_search_threshold(self, param, target_percent, total_no):
        """model/DeepFMs.py:807-823 (bisection on the magnitude threshold).  A CUDA tensor is bisected on the device in one
        cooperative launch (``dfw_prune_threshold``: same fp64 steps, same return value, one host read instead of up to 101);
        a CPU tensor takes the reference's loop as is (host-side model surgery, not the forward path)."""
        if isinstance(param, torch.Tensor) and param.is_cuda:
            thr, _ = self._device_threshold([param.detach()], float(target_percent), int(total_no))
            return float(thr.item())
        l, r = 0., 1e2
        cnt = 0
        mid = 0.
        while l < r:
            cnt += 1
            mid = (l + r) / 2
            sparse_rate = (abs(param) < mid).sum().item() * 1.0 / total_no
            if abs(sparse_rate - target_percent) < 0.0001:
                return mid
            elif sparse_rate > target_percent:
                r = mid
            else:
                l = mid
            if cnt > 100:
                break
        return mid

    # ------------------------------------------------------------------ one-shot pruning on the device
    @staticmethod
    def _spans(tensors):
        arr = (_lib.PruneSpan * len(tensors))()
        for i, t in enumerate(tensors):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("pruning needs contiguous fp32 tensors")
            arr[i].ptr, arr[i].count = t.data_ptr(), t.numel()
        return arr

    def _device_threshold(self, tensors, target, total, sym_F=0):
        """(threshold, probes) as device tensors (fp64, int32) for the concatenation of `tensors`; no host synchronisation."""
        lib = _lib.load()
        dev = tensors[0].device
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        ws = torch.empty(lib.dfw_prune_workspace_bytes(), dtype=torch.uint8, device=dev)
        thr = torch.zeros(1, dtype=torch.float64, device=dev)
        probes = torch.zeros(1, dtype=torch.int32, device=dev)
        spans = self._spans(tensors)
        with torch.cuda.device(dev):
            _lib.check(lib.dfw_prune_threshold(spans, len(tensors), sym_F, target, total, ws.data_ptr(), ws.numel(),
                                               thr.data_ptr(), probes.data_ptr(), _stream_ptr(dev)), "dfw_prune_threshold")
        return thr, probes

    def prune_one_shot(self, sparse=0.90, emb_r=1.0, emb_corr=1.0, prune_fm=1, prune_r=1, prune_deep=1):
        """The pruning block of the reference's ``fit`` (model/DeepFMs.py:647-673) applied once at the full target rates, on
        the device: one global threshold at rate ``sparse * emb_r`` over the concatenated ``fm_2nd_embeddings``, a per-tensor
        threshold at rate ``sparse`` for every parameter whose name contains ``linear`` and ``weight`` (the MLP layers and
        ``fwfm_linear``), and the symmetric mask ``|0.5 (R + R^T)| < t`` at rate ``sparse * emb_corr`` on ``field_cov.weight``.
        Every threshold is bisected by one cooperative launch and applied by one more; nothing synchronises with the host until
        the returned report is read.  Parameters are modified in place (zeros inside dense tensors, exactly what the
        reference saves) and the derived images -- pair list of the pruned R, bf16 / CSR weight images -- are rebuilt on the
        next forward.  Returns {name: (threshold, probes, zeroed)} with ``emb`` for the embedding set."""
        lib = _lib.load()
        dev = self.bias.device
        if dev.type != "cuda":
            raise RuntimeError("prune_one_shot runs on an sm_100 CUDA device (no CPU fallback); move the module first")
        jobs = []          # (report name, tensors, sym_F, rate, total)
        if prune_fm:
            embs = [p_.data for n, p_ in self.named_parameters() if "fm_2nd_embeddings" in n]
            total = sum(t.numel() for t in embs)
            jobs.append(("emb", embs, 0, sparse * emb_r, total))
        for n, p_ in self.named_parameters():
            if "linear" in n and "weight" in n and prune_deep:
                jobs.append((n, [p_.data], 0, sparse, p_.numel()))
            if n == "field_cov.weight" and prune_r:
                jobs.append((n, [p_.data], self.field_size, sparse * emb_corr, p_.numel()))
        zeroed = torch.zeros(len(jobs), dtype=torch.int64, device=dev)
        pending = []
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            for i, (name, tensors, sym_F, rate, total) in enumerate(jobs):
                thr, probes = self._device_threshold(tensors, float(rate), int(total), sym_F)
                _lib.check(lib.dfw_prune_apply(self._spans(tensors), len(tensors), sym_F, thr.data_ptr(),
                                               zeroed[i:i + 1].data_ptr(), st), "dfw_prune_apply")
                pending.append((name, thr, probes))
        self.repack()
        z = zeroed.cpu()
        return {name: (float(thr.item()), int(probes.item()), int(z[i])) for i, (name, thr, probes) in enumerate(pending)}

    def predict(self, Xi, Xv):
        return self.predict_proba(Xi, Xv) > 0.5

    def predict_proba(self, Xi, Xv):
        self.eval()
        Xi = np.array(Xi).reshape((-1, self.field_size - self.num, 1))
        return self.predict_proba_host(Xi, np.asarray(Xv, dtype=np.float32))

    def inner_predict(self, Xi, Xv):
        return self.inner_predict_proba(Xi, Xv) > 0.5

    def inner_predict_proba(self, Xi, Xv):
        self.eval()
        _, prob = self.forward(Xi, Xv, return_prob=True)
        return prob.cpu().numpy()

    def evaluate(self, Xi, Xv, y):
        return self.eval_metric(y.cpu().data.numpy(), self.inner_predict_proba(Xi, Xv))

    def print_size_of_model(self):
        n = sum(p.numel() for p in self.parameters())
        nz = sum(int((p != 0).sum()) for p in self.parameters())
        self.logger.info("parameters: %d (non-zero %d), %.2f MB fp32", n, nz, n * 4 / 1e6)
        return n, nz

    def fit(self, *a, **kw):
        raise NotImplementedError("training (fit / pruning schedule / KD) is outside the B200 forward hot path; "
                                  "train with the reference and load its state_dict here")

    def __getstate__(self):
        d = self.__dict__.copy()
        d['_plan'] = None
        d.pop('_host_ws', None)
        if 'logger' in d:
            d['logger'] = d['logger'].name
        return d

    def __setstate__(self, d):
        if 'logger' in d:
            d['logger'] = logging.getLogger(d['logger'])
        self.__dict__.update(d)


# ---------------------------------------------------------------------- standalone QR lookup
def _single_table_lookup(table: QREmbeddingBag, idx: torch.Tensor) -> torch.Tensor:
    """Run the fused gather kernel on a one-field problem: idx (B,1) -> (B, d) rows of a QR table."""
    lib = _lib.load()
    dev = table.weight_q.device
    if dev.type != "cuda":
        raise RuntimeError("QREmbeddingBag lookup needs the table on an sm_100 CUDA device (no CPU fallback)")
    _lib.require_device(dev.index)
    K = table.embedding_dim[0]
    d = (_lib.FieldDesc * 1)()
    d[0].w2, d[0].w2_r = table.weight_q.data_ptr(), table.weight_r.data_ptr()
    d[0].rows, d[0].collisions = table.num_categories, table.num_collisions
    d[0].qr_op = _lib.TABLE_QR_MULT if table.operation == "mult" else _lib.TABLE_QR_ADD
    fields = torch.frombuffer(bytearray(bytes(d)), dtype=torch.uint8).to(dev)
    zeros = torch.zeros(K + 1, dtype=torch.float32, device=dev)      # bias word + a zero fwlw row
    m = _lib.Model()
    m.struct_bytes, m.abi_version = C.sizeof(_lib.Model), _lib.DFW_ABI_VERSION
    m.field_size, m.numerical, m.embedding_size, m.flags = 1, 0, K, _lib.USE_FWLW
    m.fields, m.bias, m.fwfm_linear = fields.data_ptr(), zeros.data_ptr(), zeros[1:].data_ptr()
    B = idx.shape[0]
    idx = idx.to(torch.int64)
    E = torch.empty(B, K, dtype=torch.float32, device=dev)
    shallow = torch.empty(max(B, 1), dtype=torch.float32, device=dev)
    img = torch.zeros(lib.dfw_shallow_image_bytes(C.byref(m)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.dfw_pack_shallow(C.byref(m), img.data_ptr(), _stream_ptr(dev)), "dfw_pack_shallow")
        m.shallow_image = img.data_ptr()
        rc = lib.dfw_embed_fwfm(C.byref(m), idx.data_ptr(), idx.stride(0), 0, None, 0, 0, B, E.data_ptr(), K,
                                None, 0, shallow.data_ptr(), None, _stream_ptr(dev))
    _lib.check(rc, "dfw_embed_fwfm")
    return E
