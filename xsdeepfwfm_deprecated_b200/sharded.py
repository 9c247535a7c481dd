"""Row-sharded embedding tables across the GPUs of one node.

The reference has no distributed code at all; this is the multi-GPU form of its forward
hot path that BASELINE.json's north_star asks for: the tables are the part that shards, the MLP stays
batch-data-parallel.

Partitioning (every rank holds the same small state -- MLP, field_cov, fwfm_linear, numeric and small tables --
and a 1/P slice of every large table):

    row i of a sharded table lives on rank  i mod P  at local row  i div P        (QR tables: the quotient table is
    sharded on the quotient row, the c-row remainder table is replicated)

Three exchange modes produce bit-identical logits (rows are copied, never summed):

``p2p`` (one launch; the default, and what ``bench.py --gpus N`` runs)
    Every rank's shard is cudaMalloc'ed by ``dfw_shard_alloc`` and exported with CUDA IPC; each rank maps all peers'
    shards and puts the P pointers into the field descriptors.  The SAME fused kernel then loads a row from whichever GPU
    owns it -- over NVLink 5 / NVSwitch when it is a peer -- straight into the registers of its gather warps.  Gather,
    exchange, FwFM interaction and MLP are one kernel: no index exchange, no staging buffers, no collective on the data
    path (1260 M samples/s on 8 GPUs).

``p2p_pull`` (an option; measured slower than ``p2p`` at 2 and 8 GPUs; needs ``use_fwlw=1``)
    The same peer loads, issued by a separate small kernel (``dfw_pull_rows``) one batch AHEAD of the fused kernel: its
    128-thread CTAs fit beside a resident fused CTA, copy every sharded field's row of every sample from the owning GPU
    into a batch-ordered staging buffer in local HBM and rewrite the index columns; the fused kernel then gathers from
    local memory at single-GPU speed instead of holding an SM idle through NVLink round trips.  One ``PullLane``
    (staging + descriptors) per concurrent stream.

``nccl`` (the baseline the p2p kernels are measured against; needs ``use_fwlw=1``; it reads the split sizes back to the
    host and synchronises the stream every forward, so its number is a baseline, not a tuned collective path)
    The textbook exchange: route indices to their owners with ``all_to_all_single``, owners gather the requested rows
    (``dfw_gather_rows``), a second ``all_to_all_single`` returns them, and the fused kernel consumes them as a
    per-batch table.

``plan_exchange`` / ``route_indices`` / ``assemble_rows`` are pure torch (any device, any backend) so the routing
logic is unit-tested on CPU with gloo at world_size 2.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .model.DeepFMs import DeepFMs, _stream_ptr
from .model.QREmbeddingBag import QREmbeddingBag


# ------------------------------------------------------------------------------------------------ pure routing logic
def owner_and_local(idx: torch.Tensor, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """owner rank and local row of global row ids under the  i mod P / i div P  partition."""
    return idx % world, torch.div(idx, world, rounding_mode="floor")


def local_rows(rows: int, rank: int, world: int) -> int:
    """number of rows of a `rows`-row table that rank owns."""
    return (rows - rank + world - 1) // world if rows > rank else 0


def pull_plan(idx: torch.Tensor, collisions: int, world: int):
    """What ``dfw_pull_rows`` does with one sharded column (pure torch restatement, used by the tests): for category ids
    ``idx`` (B,) of a table with QR collision count ``collisions`` (1 = plain) it returns

        owner, local   rank that stores the row and its position in that rank's shard (row = idx // c;  row mod P, row div P)
        rewritten      the index the fused kernel gets instead: ``b * c + idx mod c`` -- row b of the (B * c)-category staging
                       table whose quotient rows are the pulled rows, same remainder as the original id
    """
    c = max(int(collisions), 1)
    row = torch.div(idx, c, rounding_mode="floor")
    owner, local = owner_and_local(row, world)
    b = torch.arange(idx.numel(), dtype=idx.dtype, device=idx.device)
    return owner, local, b * c + (idx - row * c)


def route_indices(idx: torch.Tensor, world: int):
    """Sort the requests of one rank by owner.  Returns (sorted global ids, send counts per owner, inverse permutation)."""
    owner = idx % world
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world)
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(order.numel(), device=idx.device)
    return idx[order], counts, inverse


def exchange_rows(idx: torch.Tensor, gather_fn, width: int, group=None, dtype=torch.float32) -> torch.Tensor:
    """All-to-all row exchange for ONE sharded table: idx (n,) global ids wanted by this rank -> (n, width) rows.

    gather_fn(req_ids) returns the rows of the requested global ids, all of which this rank owns.
    """
    world = dist.get_world_size(group)
    sorted_ids, send_counts, inverse = route_indices(idx, world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_l, recv_l = send_counts.tolist(), recv_counts.tolist()
    req = torch.empty(int(sum(recv_l)), dtype=idx.dtype, device=idx.device)
    dist.all_to_all_single(req, sorted_ids, output_split_sizes=recv_l, input_split_sizes=send_l, group=group)
    rows = gather_fn(req)                                                  # (sum(recv), width): rows I own
    back = torch.empty(idx.numel(), width, dtype=dtype, device=idx.device)
    dist.all_to_all_single(back, rows, output_split_sizes=send_l, input_split_sizes=recv_l, group=group)
    return back[inverse]


# ------------------------------------------------------------------------------------------------ the sharded module
class ShardedDeepFMs(DeepFMs):
    """``DeepFMs`` whose large second-order tables are row-sharded over the ranks of ``process_group``.

    Build it like ``DeepFMs`` (every rank constructs / loads the FULL state_dict, e.g. from a reference checkpoint),
    move it to this rank's GPU, then call ``shard_()``: tables with more than ``shard_threshold`` rows are cut down to
    this rank's slice and the full copies are released.  ``forward`` takes this rank's slice of the batch.
    """

    def __init__(self, *args, process_group=None, shard_threshold: int = 200, exchange: str = "p2p", **kw):
        super().__init__(*args, **kw)
        if exchange not in ("p2p", "p2p_pull", "nccl"):
            raise ValueError("exchange must be 'p2p', 'p2p_pull' or 'nccl'")
        if exchange != "p2p" and not self.use_fwlw:
            # 'nccl' and 'p2p_pull' rewrite the sharded index columns to positions in a per-batch staging table
            # (b * c + idx mod c); the first-order tables (use_fwlw=0: fm_1st_embeddings[f][idx], model/DeepFMs.py:300-309) are
            # looked up with the same column and would read row b instead of the category's row.  'p2p' keeps the ids.
            raise ValueError(f"exchange={exchange!r} needs use_fwlw=1 (the first-order tables are indexed by the original "
                             "category ids, which this exchange rewrites); use exchange='p2p'")
        self.process_group = process_group
        self.shard_threshold = shard_threshold
        self.exchange = exchange
        self._shards: Dict[int, dict] = {}          # field -> dict(ptrs=[...], own=ptr, rows=..., local_rows=...)
        self._owned_allocs: List[int] = []
        self._peer_maps: List[int] = []
        self._lanes: Dict[tuple, "PullLane"] = {}

    # -- sharding ---------------------------------------------------------------------------------------------
    def shard_(self):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        lib = _lib.load()
        group = self.process_group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = self.bias.device
        if dev.type != "cuda":
            raise RuntimeError("shard_() needs the module on this rank's CUDA device (no CPU fallback)")
        if world > _lib.DFW_MAX_RANKS:
            raise ValueError(f"at most {_lib.DFW_MAX_RANKS} ranks")
        K = self.embedding_size
        st = _stream_ptr(dev)
        handles = []
        fields = []
        with torch.cuda.device(dev):
            for f in range(self.num, self.field_size):
                emb = self.fm_2nd_embeddings[f]
                qr = isinstance(emb, QREmbeddingBag)
                param = emb.weight_q if qr else emb.weight
                rows = param.shape[0]
                if world == 1 or rows <= self.shard_threshold:
                    continue
                nloc = local_rows(rows, rank, world)
                ptr = C.c_void_p()
                _lib.check(lib.dfw_shard_alloc(max(nloc, 1) * K * 4, C.byref(ptr)), "dfw_shard_alloc")
                self._owned_allocs.append(ptr.value)
                _lib.check(lib.dfw_shard_rows(param.data_ptr(), rows, K, rank, world, ptr.value, st), "dfw_shard_rows")
                h = C.create_string_buffer(64)
                _lib.check(lib.dfw_ipc_export(ptr.value, h), "dfw_ipc_export")
                handles.append(torch.frombuffer(bytearray(h.raw), dtype=torch.uint8).clone())
                fields.append((f, rows, nloc, ptr.value, param))
            torch.cuda.synchronize(dev)
            if fields:
                mine = torch.stack(handles).to(dev)                                  # (n_sharded, 64)
                allh = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(allh, mine, group=group)
                allh = [h.cpu() for h in allh]
                for j, (f, rows, nloc, own, param) in enumerate(fields):
                    ptrs = []
                    for r in range(world):
                        if r == rank:
                            ptrs.append(own)
                            continue
                        peer = C.c_void_p()
                        raw = bytes(allh[r][j].numpy().tobytes())
                        _lib.check(lib.dfw_ipc_import(raw, C.byref(peer)), "dfw_ipc_import")
                        self._peer_maps.append(peer.value)
                        ptrs.append(peer.value)
                    self._shards[f] = dict(ptrs=ptrs, own=own, rows=rows, local_rows=nloc)
                    # release the full copy; the parameter keeps its name with zero rows
                    param.data = torch.empty(0, K, dtype=torch.float32, device=dev)
            dist.barrier(group=group)
        self.repack()
        return self

    def _patch_field_descs(self, descs, plan):
        if self.exchange not in ("p2p", "p2p_pull"):
            return
        for f, s in self._shards.items():
            d = descs[f]
            d.n_ranks = len(s["ptrs"])
            d.w2 = s["own"]
            for r, p in enumerate(s["ptrs"]):
                d.w2_shard[r] = p

    # -- forward ----------------------------------------------------------------------------------------------
    def forward(self, Xi, Xv, return_prob: bool = False):
        if self.exchange == "p2p" or not self._shards:
            return super().forward(Xi, Xv, return_prob)
        if self.exchange == "p2p_pull":
            return self._forward_pull(Xi, Xv, return_prob)
        return self._forward_nccl(Xi, Xv, return_prob)

    # -- "p2p_pull": the exchange as its own kernel, one batch ahead of the fused kernel ---------------------------
    def pull_lane(self, B: int, key=None) -> "PullLane":
        """Staging for one in-flight batch of B samples (cached per (B, key); use one lane per concurrent stream)."""
        plan = self._get_plan()
        lane = self._lanes.get((B, key))
        if lane is None or lane.plan is not plan:
            lane = self._lanes[(B, key)] = PullLane(self, plan, B)
        return lane

    def _forward_pull(self, Xi, Xv, return_prob):
        lib = _lib.load()
        plan = self._get_plan()
        dev = plan.device
        B = Xi.shape[0]
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        prob = torch.empty(B, dtype=torch.float32, device=dev) if return_prob else None
        if B == 0:
            return (logits, prob) if return_prob else logits
        want = torch.int32 if self.index_dtype == "int32" else torch.int64
        if Xi.dtype != want or Xv.dtype != torch.float32:
            raise TypeError(f"Xi must be {want} and Xv float32")
        with torch.cuda.device(dev):
            st = _stream_ptr(dev)
            lane = self.pull_lane(B, st)
            if lane.precision != self.precision:
                lane.prepare(self, self.precision)
            lane.run(lib, Xi, Xv, logits, prob, st)
        return (logits, prob) if return_prob else logits

    def _forward_nccl(self, Xi, Xv, return_prob):
        """Baseline: index all-to-all -> owner gather -> row all-to-all -> fused kernel on per-batch tables."""
        if self.index_dtype != "int64":
            raise ValueError("the NCCL exchange baseline routes int64 indices; use exchange='p2p' with index_dtype='int32'")
        lib = _lib.load()
        group = self.process_group
        world = dist.get_world_size(group)
        dev = self.bias.device
        K, B = self.embedding_size, Xi.shape[0]
        st = _stream_ptr(dev)
        Xi2 = Xi.clone()
        arange = torch.arange(B, device=dev, dtype=torch.int64)
        staged = {}
        for f, s in self._shards.items():
            emb = self.fm_2nd_embeddings[f]
            c = emb.num_collisions if isinstance(emb, QREmbeddingBag) else 1
            ids = Xi[:, f - self.num, 0]
            row_ids = torch.div(ids, c, rounding_mode="floor") if c > 1 else ids

            def gather(req, s=s):
                out = torch.empty(req.numel(), K, dtype=torch.float32, device=dev)
                if req.numel():
                    _lib.check(lib.dfw_gather_rows(s["own"], s["local_rows"], K, req.data_ptr(), req.numel(), world,
                                                   out.data_ptr(), st), "dfw_gather_rows")
                return out

            rows = exchange_rows(row_ids.contiguous(), gather, K, group=group)           # (B, K) quotient/plain rows
            staged[f] = rows
            # the fused kernel reads the staged rows as a B-row table: index b (times c plus the remainder for QR)
            Xi2[:, f - self.num, 0] = arange * c + (ids % c if c > 1 else 0)
        return self._forward_with_tables(Xi2, Xv, staged, B, return_prob)

    def _forward_with_tables(self, Xi2, Xv, staged, B, return_prob):
        lib = _lib.load()
        plan = self._get_plan()
        dev = plan.device
        F = self.field_size
        descs = (_lib.FieldDesc * F).from_buffer_copy(bytes(plan.fields_dev.cpu().numpy().tobytes()))
        for f, rows in staged.items():
            emb = self.fm_2nd_embeddings[f]
            c = emb.num_collisions if isinstance(emb, QREmbeddingBag) else 1
            descs[f].w2 = rows.data_ptr()
            descs[f].rows = B * c
            descs[f].n_ranks = 0
        fields = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        m = _lib.Model.from_buffer_copy(bytes(plan.model))
        m.fields = fields.data_ptr()
        m.shallow_image = None                       # rebuilt from the temporary descriptors inside dfw_forward
        prec = _lib.PRECISIONS[self.precision]
        plan.ensure_image(self, self.precision)
        for l in range(self.h_depth if self.use_deep else 0):
            m.Wbf16[l] = plan.model.Wbf16[l]
            m.Wbf16_lo[l] = plan.model.Wbf16_lo[l]
            m.csr[l] = plan.model.csr[l]
        ref = C.byref(m)
        nbytes = lib.dfw_forward_workspace_bytes(ref, B, prec)
        ws = torch.zeros(nbytes + 4096, dtype=torch.uint8, device=dev)
        logits = torch.empty(B, dtype=torch.float32, device=dev)
        prob = torch.empty(B, dtype=torch.float32, device=dev) if return_prob else None
        with torch.cuda.device(dev):
            rc = lib.dfw_forward(ref, Xi2.data_ptr(), Xi2.stride(0), Xi2.stride(1), Xv.data_ptr() if self.num else None,
                                 Xv.stride(0), Xv.stride(1), B, prec, ws.data_ptr(), ws.numel(), logits.data_ptr(),
                                 prob.data_ptr() if prob is not None else None, None, _stream_ptr(dev))
        _lib.check(rc, "dfw_forward")
        torch.cuda.current_stream(dev).synchronize()      # `staged`, `fields`, `ws` must outlive the kernels
        return (logits, prob) if return_prob else logits

    def release(self):
        """Unmap peers and free this rank's shards (call on every rank before exit)."""
        lib = _lib.load()
        self._plan = None
        self._lanes = {}
        for p in self._peer_maps:
            lib.dfw_ipc_close(p)
        self._peer_maps = []
        if dist.is_initialized():
            dist.barrier(group=self.process_group)
        for p in self._owned_allocs:
            lib.dfw_shard_free(p)
        self._owned_allocs = []
        self._shards = {}


class PullLane:
    """One in-flight batch of the ``p2p_pull`` exchange: the staging buffer the row-pull kernel fills, the rewritten index
    matrix, and a view of the model whose sharded fields read that staging buffer as B-row tables (own descriptors and
    shallow image; weights and images shared with the module's plan)."""

    def __init__(self, owner: ShardedDeepFMs, plan, B: int):
        lib = _lib.load()
        dev = plan.device
        self.plan, self.B = plan, B
        K, F, num = owner.embedding_size, owner.field_size, owner.num
        self.fields_sharded = sorted(owner._shards)
        n_sf = len(self.fields_sharded)
        self.sf = (C.c_int32 * n_sf)(*self.fields_sharded)
        self.staged = torch.empty(n_sf, B, K, dtype=torch.float32, device=dev)
        idt = torch.int32 if owner.index_dtype == "int32" else torch.int64
        self.xi2 = torch.empty(B, F - num, dtype=idt, device=dev)
        descs = (_lib.FieldDesc * F).from_buffer_copy(bytes(plan.fields_dev.cpu().numpy().tobytes()))
        for j, f in enumerate(self.fields_sharded):
            emb = owner.fm_2nd_embeddings[f]
            c = emb.num_collisions if isinstance(emb, QREmbeddingBag) else 1
            descs[f].w2 = self.staged[j].data_ptr()
            descs[f].rows = B * c
            descs[f].n_ranks = 0
        self.fields = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
        m = _lib.Model.from_buffer_copy(bytes(plan.model))
        m.fields = self.fields.data_ptr()
        m.shallow_image = None
        self.model = m
        self.model_ref = C.byref(m)
        self.image = torch.zeros(lib.dfw_shallow_image_bytes(self.model_ref), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.dfw_pack_shallow(self.model_ref, self.image.data_ptr(), _stream_ptr(dev)), "dfw_pack_shallow")
        m.shallow_image = self.image.data_ptr()
        self.precision = None
        self.ws = None

    def prepare(self, owner, precision: str):
        """Bind the weight images of `precision` (built by the plan) and size the forward workspace."""
        lib = _lib.load()
        self.plan.ensure_image(owner, precision)
        pm, m = self.plan.model, self.model
        for l in range(owner.h_depth if owner.use_deep else 0):
            m.Wbf16[l], m.Wbf16_lo[l], m.csr[l] = pm.Wbf16[l], pm.Wbf16_lo[l], pm.csr[l]
        self.precision = precision
        self.prec = _lib.PRECISIONS[precision]
        n = lib.dfw_forward_workspace_bytes(self.model_ref, self.B, self.prec)
        self.ws = torch.zeros(n + 4096, dtype=torch.uint8, device=self.plan.device)
        return self

    def enqueue_pull(self, lib, xi_ptr, xi_sb, xi_sc, pull_stream):
        """Row pull of one batch into this lane's staging buffer (``dfw_pull_rows``)."""
        rc = lib.dfw_pull_rows(self.plan.model_ref, self.sf, len(self.fields_sharded), xi_ptr, xi_sb, xi_sc, self.B,
                               self.staged.data_ptr(), self.xi2.data_ptr(), None, pull_stream)
        if rc:
            _lib.check(rc, "dfw_pull_rows")

    def enqueue_forward(self, lib, xv_ptr, xv_sb, xv_sc, logits_ptr, prob_ptr, stream):
        """Fused forward on the staged rows; the caller orders it after the pull when the streams differ."""
        Cc = self.xi2.shape[1]
        rc = lib.dfw_forward(self.model_ref, self.xi2.data_ptr(), Cc, 1, xv_ptr, xv_sb, xv_sc, self.B, self.prec,
                             self.ws.data_ptr(), self.ws.numel(), logits_ptr, prob_ptr, None, stream)
        if rc:
            _lib.check(rc, "dfw_forward")

    def enqueue(self, lib, xi_ptr, xi_sb, xi_sc, xv_ptr, xv_sb, xv_sc, logits_ptr, prob_ptr, pull_stream, stream):
        self.enqueue_pull(lib, xi_ptr, xi_sb, xi_sc, pull_stream)
        self.enqueue_forward(lib, xv_ptr, xv_sb, xv_sc, logits_ptr, prob_ptr, stream)

    def run(self, lib, Xi, Xv, logits, prob, st):
        self.enqueue(lib, Xi.data_ptr(), Xi.stride(0), Xi.stride(1), Xv.data_ptr() if Xv.numel() else None, Xv.stride(0),
                     Xv.stride(1), logits.data_ptr(), prob.data_ptr() if prob is not None else None, st, st)
