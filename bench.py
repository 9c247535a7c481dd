#!/usr/bin/env python
"""Benchmark of the DeepFwFM forward hot path (BASELINE.json metric: inference samples/s + roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|bf16|fp32_csr]

A step = one forward of one synthetic Criteo-shaped batch (BASELINE config 2: DeepFwFM dense, fwfm+deep+fwlw,
MLP 400x400x400, paper-Criteo cardinalities, B = 4096 per GPU).  Prints ONE JSON line (rank 0).

  value     samples/s with inputs resident in HBM: K steps replayed from CUDA graphs over distinct batches,
            CUDA events on the launch stream, max over ranks
  e2e       samples/s through the host-buffer C-ABI call dfw_forward_host: pinned host batches -> H2D -> kernels
            -> sigmoid -> D2H -> sync, every step inside the timed region
  roofline  the dominant kernel stage, timed live with CUDA events (algorithmic bytes / flops per SURVEY 8(d))
  parity    before anything is timed, the timed call path's logits for this rank's batch 0 against oracle/closed_form.py
            on the model's real weights (every --gpus N, every --workload)
  cpu_baseline / --impl reference: the UNMODIFIED reference module from baseline/_ref (a git-ignored copy of
            /root/reference/{model,utils} made by __graft_entry__.build(); kind "reference"), fp32 torch CPU on the host
            cores; oracle/torch_port.py (kind "port") only when that copy is absent
  reference_cuda  supplementary: the same unmodified module with use_cuda=True on this GPU (eager PyTorch)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "DeepFwFM inference samples/sec"
UNIT = "samples/s"
K_EMB, NODES, DEPTH = 10, 400, 3
SHARD_ROWS = int(os.environ.get("DFW_BENCH_SHARD_ROWS", "65536"))
FIELD = NUM = CATS = 0
SIZES, MODEL_KW, WORKLOAD_TEXT, XV_UNIT, PRUNED = None, {}, "", False, False
ALG_BYTES_PER_SAMPLE = ALG_BYTES_PER_BATCH = MLP_FLOPS_PER_SAMPLE = 0


def set_workload(name):
    """criteo = BASELINE config 2 (the headline); criteo_qr = config 4(ii); twitter = config 5's shape (supplementary lines)."""
    global FIELD, NUM, CATS, SIZES, MODEL_KW, WORKLOAD_TEXT, XV_UNIT, ALG_BYTES_PER_SAMPLE, ALG_BYTES_PER_BATCH, MLP_FLOPS_PER_SAMPLE
    from xsdeepfwfm_deprecated_b200.utils import workloads as synth
    global PRUNED
    PRUNED = False
    XV_UNIT = False
    if name == "criteo":
        FIELD, NUM, SIZES, MODEL_KW = 39, 13, synth.CRITEO_PAPER, {}
        WORKLOAD_TEXT = ("BASELINE config 2: DeepFwFM dense (fwfm+deep+fwlw), F=39 (13 numeric), K=10, MLP 400x400x400, "
                         "paper-Criteo cardinalities (1.33 M rows, 53 MB fp32), uniform indices")
    elif name == "criteo_pruned":
        FIELD, NUM, SIZES, MODEL_KW, PRUNED = 39, 13, synth.CRITEO_PAPER, {}, True
        WORKLOAD_TEXT = ("BASELINE config 3: config 2's model after the reference's one-shot pruning recipe (DeepFMs.prune_one_shot: "
                         "sparse 0.9 on the MLP layers, fwfm_linear and |R|; 0.9*0.444 over the stacked embeddings), zeros kept inside "
                         "the dense tensors as the reference saves them; uniform indices")
    elif name == "criteo_qr":
        FIELD, NUM, SIZES = 39, 13, synth.CRITEO_KAGGLE
        MODEL_KW = dict(embedding_bag=1, qr_flag=1, qr_operation="mult", qr_collisions=4, qr_threshold=200)
        WORKLOAD_TEXT = ("BASELINE config 4(ii): DeepFwFM with QREmbeddingBag (mult, c=4, threshold 200) on the un-thresholded "
                         "Kaggle cardinalities (33.8 M categories -> 8.44 M quotient rows, 338 MB fp32), F=39, K=10, MLP "
                         "400x400x400, uniform indices")
    elif name == "twitter":
        FIELD, NUM, SIZES, MODEL_KW, XV_UNIT = 47, 11, synth.TWITTER_SYNTH, {}, True
        WORKLOAD_TEXT = ("BASELINE config 5 shape: DeepFwFM dense (fwfm+deep+fwlw), Twitter RecSys2020 layout F=47 (11 numeric, "
                         "36 categorical), K=10, MLP 400x400x400, synthetic cardinalities (69.2 M rows, 2.77 GB fp32), uniform indices")
    else:
        raise ValueError(name)
    CATS = FIELD - NUM
    FK = FIELD * K_EMB
    mlp_params = FK * NODES + NODES + (DEPTH - 1) * (NODES * NODES + NODES) + NODES
    ALG_BYTES_PER_SAMPLE = CATS * 8 + NUM * 4 + CATS * K_EMB * 4 + 4          # Criteo: 1304, Twitter: 1776
    ALG_BYTES_PER_BATCH = mlp_params * 4 + FIELD * FIELD * 4 + FIELD * K_EMB * 4 + NUM * K_EMB * 4   # read once per launch
    MLP_FLOPS_PER_SAMPLE = 2 * (FK * NODES + (DEPTH - 1) * NODES * NODES + NODES)                    # Criteo: 952,800


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--workload", default="criteo", choices=["criteo", "criteo_pruned", "criteo_qr", "twitter"],
                    help="criteo = BASELINE config 2 (the headline line); the others are supplementary shapes")
    ap.add_argument("--precision", default=os.environ.get("DFW_BENCH_PRECISION", "bf16x3"),
                    choices=["bf16x3", "bf16", "fp32", "fp32_csr"])
    ap.add_argument("--nbatches", type=int, default=256, help="distinct input batches cycled (> L2 in total)")
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--streams", type=int, default=16,
                    help="concurrent streams the K independent forwards are spread over (1 = strictly back to back)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------- workload
def make_model(device, precision, feature_sizes, world=1, exchange=None, index_dtype="int64", keep_weights=None):
    # exchange: "p2p" (default) = peer rows fetched over NVLink by the fused kernel's own gather (one launch per step; with two tiles
    # per CTA pair the peer round trips of the second tile run under the first tile's MLP: 26.1 us/step at 2 and at 8 GPUs);
    # "p2p_pull" = the same loads by a separate kernel one batch ahead into a local staging buffer (two launches; its CTAs cannot
    # share an SM with the fused kernel's: 31.2 us/step)
    exchange = exchange or os.environ.get("DFW_BENCH_EXCHANGE", "p2p")
    from xsdeepfwfm_deprecated_b200.model import DeepFMs
    kw = dict(embedding_size=K_EMB, h_depth=DEPTH, deep_nodes=NODES, use_fm=False, use_fwfm=True, use_deep=True,
              use_fwlw=True, use_lw=False, use_cuda=True, numerical=NUM, random_seed=42, precision=precision,
              index_dtype=index_dtype, throughput_hint=True, **MODEL_KW)
    if world > 1:
        from xsdeepfwfm_deprecated_b200.sharded import ShardedDeepFMs
        # shard what is large: tables above 65,536 rows (2.6 MB at K = 10) -- 6 of the 26 paper-Criteo tables = 89 % of the table bytes,
        # 16 of the 36 Twitter tables = 99.9 %; a 12 k-row table is replicated (480 KB) rather than fetched over NVLink per sample
        m = ShardedDeepFMs(FIELD, feature_sizes, exchange=exchange, shard_threshold=SHARD_ROWS, **kw)
    else:
        m = DeepFMs(FIELD, feature_sizes, **kw)
    m = m.to(device)
    m.init_weights()                       # the reference's init distributions, on the device
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if "fm_2nd_embeddings" in n_ and not n_.endswith("weight_r"):
                p_.mul_(10.0)                              # trained-scale embeddings
    if PRUNED:              # config 3: the reference's one-shot recipe (model/DeepFMs.py:647-673) at the paper's rates, on the device
        m.eval()
        m.prune_one_shot(sparse=0.9, emb_r=0.444, emb_corr=1.0)
    if keep_weights is not None:        # fp32 parameters as the oracle takes them -- captured BEFORE the tables are sharded
        keep_weights.update({k: v.detach().cpu().numpy() for k, v in m.state_dict().items()})
    if world > 1:
        m.shard_()          # every rank built identical full tables (same seed); keep only this rank's rows
    return m.eval().freeze()


PARITY_BOUND = {"fp32": 1e-5, "bf16x3": 1e-5, "fp32_csr": 1e-5, "bf16": 5e-4}


def oracle_parity(weights, Xi0, Xv0, got, precision, n=256):
    """Every timed arm checks itself first: the logits the timed call path produced for the first `n` samples of this rank's
    batch 0 against the fp64 closed form (oracle/closed_form.py, pinned to the reference by tests/golden) evaluated on the
    model's real fp32 parameters (for sharded tables: captured before shard_()).  The oracle is the checker here, nothing else."""
    from oracle import closed_form
    from oracle.config import PathConfig
    cfg = PathConfig(FIELD, SIZES, embedding_size=K_EMB, numerical=NUM, use_fm=False, use_fwfm=True, use_deep=True,
                     use_fwlw=True, h_depth=DEPTH, deep_nodes=NODES, **MODEL_KW)
    n = min(n, Xi0.shape[0])
    ref = closed_form.forward(cfg, weights, Xi0[:n].cpu().numpy(), Xv0[:n].cpu().numpy())
    g = got[:n].double().cpu().numpy()
    scale = float(np.abs(ref["logit"]).max())
    rel = float(np.abs(g - ref["logit"]).max() / scale)
    shallow = ref["first"] + ref["second"] + float(weights["bias"][0])
    deep_rel = float(np.abs((g - shallow) - ref["deep"]).max() / max(float(np.abs(ref["deep"]).max()), 1e-30))
    bound = PARITY_BOUND[precision]
    out = dict(max_rel=rel, bound=bound, n=int(n), vs="oracle/closed_form (fp64)", max_abs_logit=round(scale, 4),
               deep_term_rel=deep_rel, ok=bool(rel <= bound))
    assert rel <= bound, f"parity: max|dlogit|/max|logit| = {rel:.3e} > {bound:g} on the timed path"
    return out


def make_batches(device, feature_sizes, B, nb, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cats = torch.tensor(feature_sizes[NUM:], device=device, dtype=torch.float64)
    u = torch.rand(nb, B, FIELD - NUM, generator=g, device=device, dtype=torch.float64)
    Xi = (u * cats).long().clamp_(max=int(max(feature_sizes)) - 1)
    Xi = torch.minimum(Xi, (cats - 1).long()).unsqueeze(-1).contiguous()        # (nb, B, 26, 1) uniform indices
    if XV_UNIT:     # Twitter: MinMax-scaled dense features (data/large/preprocess_twitter.py:102-103)
        Xv = torch.rand(nb, B, NUM, generator=g, device=device)
    else:
        Xv = torch.randint(0, 50, (nb, B, NUM), generator=g, device=device).float()
    return Xi, Xv


def time_events(fn, iters, stream):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(iters):
        fn(i)
    b.record(stream)
    b.synchronize()
    return a.elapsed_time(b) / iters      # ms per call


def run_ours(args):
    from oracle import synth
    from xsdeepfwfm_deprecated_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    sizes = SIZES
    B, nb = args.batch, args.nbatches
    weights = {}
    model = make_model(device, args.precision, sizes, world, keep_weights=weights)
    Xi, Xv = make_batches(device, sizes, B, nb, seed=rank)
    plan = model._get_plan()
    plan.ensure_image(model, args.precision)
    prec = _lib.PRECISIONS[args.precision]
    ws = plan.get_workspace(lib.dfw_forward_workspace_bytes(plan.model_ref, B, prec))
    logits = torch.empty(nb, B, device=device)
    stream = torch.cuda.Stream(device)
    sp = stream.cuda_stream

    # The K steps are K independent forwards (one batch each).  They are spread round-robin over `--streams` streams -- the
    # way a server runs concurrent requests: a forward occupies 128 of the 148 SMs with one CTA each, so the next forward's
    # CTAs start on the idle SMs and on every SM the previous one frees, instead of waiting for its slowest CTA.
    nstreams = max(1, args.streams)
    side = [torch.cuda.Stream(device) for _ in range(nstreams - 1)]
    lanes = [stream] + side
    wss = [ws] + [torch.zeros_like(ws) for _ in side]

    # sharded tables: the exchange runs as its own small kernel (dfw_pull_rows: direct peer loads over NVLink into a staging
    # buffer) in front of the fused kernel of the same stream; with several streams the pull of one batch runs under the
    # fused kernels of the others.  One PullLane (staging + descriptors) per stream.
    pull = ([model.pull_lane(B, k).prepare(model, args.precision) for k in range(nstreams)]
            if world > 1 and model._shards and model.exchange == "p2p_pull" else None)
    # The pulls go to ONE high-priority stream: the block scheduler issues a later kernel's CTAs only when every earlier kernel of
    # the same priority has none left to issue, so at equal priority a pull would queue behind the pending CTAs of the other
    # lanes' fused kernels instead of running beside them.
    pull_stream = torch.cuda.Stream(device, priority=-1) if pull else None
    ev_pulled = [torch.cuda.Event() for _ in range(nstreams)] if pull else None
    ev_used = [torch.cuda.Event() for _ in range(nstreams)] if pull else None
    lane_used = [False] * nstreams

    def step(i, lane=0):
        j = i % nb
        if pull:
            if lane_used[lane]:
                pull_stream.wait_event(ev_used[lane])          # the lane's staging buffer has been consumed
            pull[lane].enqueue_pull(lib, Xi[j].data_ptr(), CATS, 1, pull_stream.cuda_stream)
            ev_pulled[lane].record(pull_stream)
            lanes[lane].wait_event(ev_pulled[lane])
            pull[lane].enqueue_forward(lib, Xv[j].data_ptr(), NUM, 1, logits[j].data_ptr(), None, lanes[lane].cuda_stream)
            ev_used[lane].record(lanes[lane])
            lane_used[lane] = True
            return
        rc = lib.dfw_forward(plan.model_ref, Xi[j].data_ptr(), CATS, 1, Xv[j].data_ptr(), NUM, 1, B, prec,
                             wss[lane].data_ptr(), wss[lane].numel(), logits[j].data_ptr(), None, None,
                             lanes[lane].cuda_stream)
        if rc:
            _lib.check(rc, "dfw_forward")

    def steps_on_lanes(first, count):
        """`count` forwards starting at batch `first`, round-robin over the streams, forked from and joined into `stream`."""
        for sd in side:
            sd.wait_stream(stream)
        if pull:
            pull_stream.wait_stream(stream)
            for k in range(nstreams):
                lane_used[k] = False       # events of an earlier segment / capture are not waited on (the join below orders them)
        for i in range(count):
            step(first + i, i % nstreams)
        for sd in side:
            stream.wait_stream(sd)
        if pull:
            stream.wait_stream(pull_stream)

    def barrier():
        torch.cuda.synchronize(device)
        if dist:
            dist.barrier()
        torch.cuda.synchronize(device)

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3)):
            step(i)
    barrier()
    l0 = lib.dfw_launch_count()
    step(0)
    launches_per_step = lib.dfw_launch_count() - l0
    torch.cuda.synchronize(device)
    # parity of the timed call path (this rank's batch 0 as step() just computed it) against the oracle, before anything is timed
    parity = oracle_parity(weights, Xi[0], Xv[0], logits[0], args.precision)
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref_cuda = reference_cuda(device, B, weights, Xi[0], Xv[0])
        if "logits" in ref_cuda:        # the reference's own fp32 GPU forward against ours on the same batch
            r_ = ref_cuda.pop("logits").float()
            ref_cuda["max_rel_vs_ours"] = float((r_ - logits[0]).abs().max() / r_.abs().max())
        torch.cuda.empty_cache()
    del weights
    if dist:        # every rank checks its own batch; the line reports the worst
        t = torch.tensor([parity["max_rel"], parity["deep_term_rel"]], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        parity.update(max_rel=float(t[0].item()), deep_term_rel=float(t[1].item()), ranks=world)

    # K steps as CUDA-graph replays of G-step segments (each segment walks distinct batches)
    graphs, tails, G = [], {}, 0
    if args.graph:
        G = min(64, args.steps)          # one graph holds every timed step when K <= 64 (no graph boundary drains the 4 streams)
        nseg = max(1, min(nb // G, 16 if G <= 16 else 4))
        for s in range(nseg):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                steps_on_lanes(s * G, G)
            graphs.append(g)

    def tail_graph(r):       # the k mod G steps left over run from a graph of their own: no eager launches in the timed region
        if r not in tails:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                steps_on_lanes((nb - r) % nb, r)
            tails[r] = g
        return tails[r]

    if graphs:
        for k_ in {max(args.warmup, 3) % G, args.steps % G} - {0}:
            tail_graph(k_)

    def run_steps(k):
        if not graphs:
            with torch.cuda.stream(stream):
                steps_on_lanes(0, k)
            return
        done, s = 0, 0
        with torch.cuda.stream(stream):
            while done + G <= k:
                graphs[s % len(graphs)].replay()
                done += G
                s += 1
            if k - done:
                tail_graph(k - done).replay()

    run_steps(max(args.warmup, 3))
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    run_steps(args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-kernel timing for the roofline (live, CUDA events, same stream, same inputs) -------------
    def graph_time(fn, n_it):
        """ms per call of `fn`, replayed from a CUDA graph of n_it calls over distinct batches (device time only)."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(n_it):
                fn(i)
        with torch.cuda.stream(stream):
            g.replay()
            torch.cuda.synchronize(device)
            return time_events(lambda i: g.replay(), 5, stream) / n_it

    pk = peaks()
    n_it = 64
    fused = bool(lib.dfw_fused_supported(plan.model_ref, prec))
    step_bytes = ALG_BYTES_PER_SAMPLE * B + ALG_BYTES_PER_BATCH
    if fused:
        # one kernel per step: gather + FwFM (HBM/L2 side) and the MLP (tensor side) overlap inside it
        def fused_only(i):          # sharded tables: the fused kernel alone, on the rows the last pull staged
            pl = pull[0]
            rc = lib.dfw_forward(pl.model_ref, pl.xi2.data_ptr(), CATS, 1, Xv[i % nb].data_ptr(), NUM, 1, B, prec, pl.ws.data_ptr(),
                                 pl.ws.numel(), logits[i % nb].data_ptr(), None, None, sp)
            if rc:
                _lib.check(rc, "dfw_forward")

        t_f = graph_time(fused_only if pull else step, n_it)
        # The fused kernel is persistent with 64 samples per CTA: one B = 4096 launch occupies 64 of the 148 SMs and is built to
        # run beside its neighbours (the timed region keeps `--streams` launches in flight).  Its roofline entry is therefore the
        # timed region's own: all launches' algorithmic FLOPs / bytes over the CUDA-event time of the K timed steps (ms =
        # ms_per_step = region time / launches); the duration of ONE launch running alone is reported beside it.
        stage = {
            "fused_forward": dict(ms=ms_per_step, bound="tensor", achieved=MLP_FLOPS_PER_SAMPLE * B / (ms_per_step * 1e-3) / 1e12,
                                  peak=pk["tf_burst"], unit="TFLOP/s"),
            "fused_forward_hbm_view": dict(ms=ms_per_step, bound="hbm", achieved=step_bytes / (ms_per_step * 1e-3) / 1e9, peak=pk["hbm"],
                                           unit="GB/s"),
            "fused_forward_one_launch_alone": dict(ms=t_f, bound="tensor", achieved=MLP_FLOPS_PER_SAMPLE * B / (t_f * 1e-3) / 1e12,
                                                   peak=pk["tf_burst"], unit="TFLOP/s"),
        }
        if world == 1 and not pull:
            # supplementary: one launch over a batch large enough to give every SM several tiles (the persistent pipeline's own rate)
            BL = 65536
            XiL, XvL = make_batches(device, sizes, BL, 2, seed=1234)
            outL = torch.empty(BL, device=device)

            def big(i):
                rc = lib.dfw_forward_fused(plan.model_ref, XiL[i % 2].data_ptr(), CATS, 1, XvL[i % 2].data_ptr(), NUM, 1, BL, prec,
                                           outL.data_ptr(), None, None, sp)
                if rc:
                    _lib.check(rc, "dfw_forward_fused")

            t_big = graph_time(big, 8)
            stage["fused_forward_batch_65536"] = dict(ms=t_big, bound="tensor", achieved=MLP_FLOPS_PER_SAMPLE * BL / (t_big * 1e-3) / 1e12,
                                                      peak=pk["tf_burst"], unit="TFLOP/s", samples_per_s=round(BL / (t_big * 1e-3), 1))
            del XiL, XvL, outL
        if pull:
            def pull_only(i):
                pl = pull[0]
                j = i % nb
                rc = lib.dfw_pull_rows(plan.model_ref, pl.sf, len(pl.fields_sharded), Xi[j].data_ptr(), CATS, 1, B,
                                       pl.staged.data_ptr(), pl.xi2.data_ptr(), None, sp)
                if rc:
                    _lib.check(rc, "dfw_pull_rows")

            t_p = graph_time(pull_only, n_it)
            n_sf = len(pull[0].fields_sharded)
            pull_bytes = B * n_sf * (K_EMB * 4 * 2 + 8 + 8) + B * (CATS - n_sf) * 16     # rows in + out, indices in + out
            stage["pull_rows"] = dict(ms=t_p, bound="hbm", achieved=pull_bytes / (t_p * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s")
        dom = "fused_forward"
    else:
        FK = FIELD * K_EMB
        ldE, ldEb = (FK + 3) // 4 * 4, (FK + 7) // 8 * 8
        Bp = (B + 127) // 128 * 128
        E = torch.zeros(Bp, ldE, device=device)
        Eb = torch.zeros(Bp, ldEb, device=device, dtype=torch.bfloat16)
        shallow = torch.zeros(Bp, device=device)
        mws = torch.zeros(lib.dfw_mlp_workspace_bytes(plan.model_ref, B, prec) + 4096, dtype=torch.uint8, device=device)
        bf = args.precision == "bf16"

        def embed(i):
            j = i % nb
            rc = lib.dfw_embed_fwfm(plan.model_ref, Xi[j].data_ptr(), CATS, 1, Xv[j].data_ptr(), NUM, 1, B,
                                    None if bf else E.data_ptr(), ldE, Eb.data_ptr() if bf else None, ldEb,
                                    shallow.data_ptr(), None, sp)
            if rc:
                _lib.check(rc, "dfw_embed_fwfm")

        mlp_fn = {"fp32": lib.dfw_mlp_fp32, "bf16x3": lib.dfw_mlp_fp32, "fp32_csr": lib.dfw_mlp_csr,
                  "bf16": lib.dfw_mlp_bf16}[args.precision]

        def mlp(i):
            rc = mlp_fn(plan.model_ref, Eb.data_ptr() if bf else E.data_ptr(), ldEb if bf else ldE, B, shallow.data_ptr(),
                        mws.data_ptr(), mws.numel(), logits[i % nb].data_ptr(), None, sp)
            if rc:
                _lib.check(rc, "dfw_mlp")

        with torch.cuda.stream(stream):
            for i in range(5):
                embed(i); mlp(i)
            torch.cuda.synchronize(device)
        t_embed = graph_time(embed, n_it)
        t_mlp = graph_time(mlp, n_it)
        embed_bytes = ALG_BYTES_PER_SAMPLE * B + FIELD * FIELD * 4 + FIELD * K_EMB * 4 + NUM * K_EMB * 4
        stage = {
            "embed_fwfm": dict(ms=t_embed, bound="hbm", achieved=embed_bytes / (t_embed * 1e-3) / 1e9, peak=pk["hbm"],
                               unit="GB/s"),
            "mlp": dict(ms=t_mlp, bound="tensor", achieved=MLP_FLOPS_PER_SAMPLE * B / (t_mlp * 1e-3) / 1e12,
                        peak=pk["tf_burst"], unit="TFLOP/s"),
        }
        dom = "mlp" if t_mlp >= t_embed else "embed_fwfm"
    d = stage[dom]
    traffic = None      # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if fused and B == 4096 and args.workload == "criteo" and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.precision)
    roofline = dict(kernel=dom, bound=d["bound"], achieved=round(d["achieved"], 3), peak=d["peak"], unit=d["unit"],
                    frac=round(d["achieved"] / d["peak"], 5), traffic=traffic, peak_source=pk["src"],
                    ms_per_launch=round(d["ms"], 5),
                    stages={k: dict(ms=round(v["ms"], 5), achieved=round(v["achieved"], 3), unit=v["unit"],
                                    frac=round(v["achieved"] / v["peak"], 5),
                                    **({"samples_per_s": v["samples_per_s"]} if "samples_per_s" in v else {})) for k, v in stage.items()},
                    whole_step_hbm_frac=round(step_bytes / (ms_per_step * 1e-3) / 1e9 / pk["hbm"], 5))

    # ---- e2e: host buffers through dfw_forward_host_stream -------------------------------------------------
    # every step's Xi/Xv start in pinned host memory and its probabilities end there: H2D -> kernel(s) -> sigmoid -> D2H per
    # batch on rotating streams (copies overlap kernels), one host synchronisation per call of `nh` steps
    nh = max(1, min(64, args.steps, nb))
    hXi = torch.empty(nh, B, CATS, dtype=torch.int64).pin_memory()
    hXv = torch.empty(nh, B, NUM, dtype=torch.float32).pin_memory()
    hXi.copy_(Xi[:nh, :, :, 0].cpu()); hXv.copy_(Xv[:nh].cpu())
    hout = torch.empty(nh, B, dtype=torch.float32).pin_memory()
    hws = torch.zeros(lib.dfw_forward_host_stream_workspace_bytes(plan.model_ref, B, prec) + 4096, dtype=torch.uint8,
                      device=device)

    def e2e_steps(k):
        done = 0
        while done < k:
            n = min(nh, k - done)
            rc = lib.dfw_forward_host_stream(plan.model_ref, hXi.data_ptr(), hXv.data_ptr(), n * B, B, prec, hws.data_ptr(),
                                             hws.numel(), None, hout.data_ptr(), sp)
            if rc:
                _lib.check(rc, "dfw_forward_host_stream")
            done += n

    mapped = bool(lib.dfw_host_transport_is_mapped(plan.model_ref, prec, hXi.data_ptr(), hXv.data_ptr(), None, hout.data_ptr()))
    e2e_steps(max(3, args.warmup, nh))      # >= nh so that every row of hout is written before the check
    # the result really is the forward of the host inputs (guards against timing a path that skips work)
    with torch.no_grad():
        chk = torch.stack([torch.sigmoid(model(Xi[j], Xv[j])) for j in range(nh)]).cpu()
    assert torch.allclose(hout, chk, atol=1e-6, rtol=1e-5), "e2e output differs from the device-resident forward"
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize(device)
    t_e2e = time.perf_counter() - t0
    if dist:
        t = torch.tensor([t_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e = dict(value=round(world * B * args.steps / t_e2e, 1), unit=UNIT,
               h2d_bytes_per_step=B * (CATS * 8 + NUM * 4), d2h_bytes_per_step=B * 4,
               ms_per_step=round(t_e2e / args.steps * 1e3, 4),
               transport="mapped" if mapped else "staged",
               api=("dfw_forward_host_stream, mapped transport: Xi/Xv travel from pinned host memory in chunks of up to 8 steps per "
                    "copy-engine transfer (1, 2, 4, then 8 steps: SM-issued PCIe reads saturate at ~34 GB/s on these boxes, a few-MB "
                    "DMA reaches 40-45) into rotating staging slots; each step is one fused launch that waits on its chunk's event and "
                    "stores its probabilities straight into pinned host memory (no D2H copy; 6 rotating compute streams" if mapped else
                    "dfw_forward_host_stream, staged transport (pinned host Xi/Xv -> cudaMemcpyAsync H2D -> forward + sigmoid "
                    "-> D2H into pinned host memory, every step; 3 rotating streams") +
                   f", one host sync per {nh} steps); timed with the host clock")
    # supplementary: the same end-to-end path fed with the packed int32 index format (DFW_XI_INT32); the
    # headline e2e above keeps the reference's int64 indices
    if world == 1:
        try:
            m32 = make_model(device, args.precision, sizes, 1, index_dtype="int32")
            p32 = m32._get_plan()
            p32.ensure_image(m32, args.precision)
            hXi32 = hXi.to(torch.int32).pin_memory()

            def e2e32(k):
                done = 0
                while done < k:
                    n = min(nh, k - done)
                    rc = lib.dfw_forward_host_stream(p32.model_ref, hXi32.data_ptr(), hXv.data_ptr(), n * B, B, prec,
                                                     hws.data_ptr(), hws.numel(), None, hout.data_ptr(), sp)
                    if rc:
                        _lib.check(rc, "dfw_forward_host_stream")
                    done += n

            ref_out = hout[0].clone()
            e2e32(max(3, args.warmup))
            assert torch.equal(hout[0], ref_out), "int32-index path differs from int64"
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            e2e32(args.steps)
            torch.cuda.synchronize(device)
            t32 = time.perf_counter() - t0
            e2e["int32_indices"] = dict(value=round(B * args.steps / t32, 1), unit=UNIT,
                                        h2d_bytes_per_step=B * (CATS * 4 + NUM * 4), ms_per_step=round(t32 / args.steps * 1e3, 4))
            del m32
        except Exception as ex:        # supplementary only: never lose the bench line over it
            e2e["int32_indices"] = {"error": str(ex)[:200]}
    clk = clocks.stop() if rank == 0 else None

    # supplementary: the public module in a plain Python loop (one ctypes call per step on one stream: no graph, no launches in flight)
    eager = None
    if rank == 0 and world == 1:
        try:
            with torch.no_grad():
                for j in range(3):
                    model(Xi[j % nb], Xv[j % nb])
                torch.cuda.synchronize(device)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                for j in range(args.steps):
                    model(Xi[j % nb], Xv[j % nb])
                eb.record()
                eb.synchronize()
            ems = ea.elapsed_time(eb) / args.steps
            eager = dict(value=round(B / (ems * 1e-3), 1), unit=UNIT, ms_per_step=round(ems, 4),
                         what="model(Xi, Xv) in a Python loop, device-resident inputs, one stream, no CUDA graph")
        except Exception as ex:        # supplementary only: never lose the bench line over it
            eager = {"error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_baseline(B, args.cpu_seconds)
    if rank == 0:
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp32_csr": "f32", "bf16": "bf16 operands, f32 accumulate (shallow part f32)",
                      "bf16x3": "f32 (MLP products as 3 split-bf16 tcgen05 MMAs with f32 accumulate, inside the fp32 "
                                "parity bound 1e-5*max|logit|; gather/FwFM in f32)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT,
                       "batch_per_gpu": B, "precision": args.precision,
                       "l2": f"inputs cycle over {nb} distinct batches ({nb * B * (CATS * 8 + NUM * 4) / 1e6:.0f} MB > 126 MB L2)" +
                             ("; the 53 MB of tables + 1.9 MB of weights stay L2-resident by design" if args.workload == "criteo"
                              else "; the tables are larger than L2 too"),
                       "launch": (f"CUDA graphs of {G} steps" if graphs else "stream launches") +
                                 (f", the independent forwards round-robin over {nstreams} concurrent streams" if nstreams > 1
                                  else ", strictly back to back on one stream"),
                       "tables": "one GPU" if world == 1 else
                                 f"{len(model._shards)} of {CATS} categorical tables (those above {SHARD_ROWS} rows) row-sharded over {world} GPUs (row i on rank i mod P), the rest replicated; " + ("rows fetched by direct peer loads over NVLink inside the fused gather kernel (no collective)"
                                    if not pull else "rows fetched by direct peer loads over NVLink (no collective) by a pull kernel "
                                    "that runs ahead of the fused kernel into a local staging buffer (exchange='p2p_pull')")},
            "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps), "clocks": clk,
            "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "reference_cuda": ref_cuda, "eager_loop": eager,
        }
        print(json.dumps(out))
    if dist:
        model.release()
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- CPU arms
REF_DIR = os.path.join(ROOT, "baseline", "_ref")        # git-ignored copy of /root/reference/{model,utils} made by build()


def _oracle_cfg():
    from oracle.config import PathConfig
    return PathConfig(FIELD, SIZES, embedding_size=K_EMB, numerical=NUM, use_fm=False, use_fwfm=True, use_deep=True,
                      use_fwlw=True, h_depth=DEPTH, deep_nodes=NODES, **MODEL_KW)


def reference_module(cfg, sd, cuda=False):
    """The UNMODIFIED reference class (model/DeepFMs.py:47) from baseline/_ref with the given state_dict, in eval mode;
    None when the copy is absent (then the arms fall back to oracle/torch_port.py and say kind = 'port')."""
    if not os.path.exists(os.path.join(REF_DIR, "model", "DeepFMs.py")):
        return None
    import logging
    import warnings
    warnings.filterwarnings("ignore")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from model.DeepFMs import DeepFMs as RefDeepFMs          # the reference itself
    log = logging.getLogger("reference")
    log.addHandler(logging.NullHandler())
    log.propagate = False
    m = RefDeepFMs(cfg.field_size, cfg.feature_sizes, embedding_size=cfg.embedding_size, h_depth=cfg.h_depth,
                   deep_nodes=cfg.deep_nodes, use_fm=cfg.use_fm, use_fwfm=cfg.use_fwfm, use_deep=cfg.use_deep,
                   use_fwlw=cfg.use_fwlw, use_lw=cfg.use_lw, use_cuda=bool(cuda), numerical=cfg.numerical,
                   embedding_bag=cfg.embedding_bag, qr_flag=cfg.qr_flag, qr_operation=cfg.qr_operation,
                   qr_collisions=cfg.qr_collisions, qr_threshold=cfg.qr_threshold, logger=log)
    m.load_state_dict(sd, strict=True)
    if cuda:            # model/DeepFMs.py:961-963
        m.use_cuda = True
        m = m.cuda()
    return m.eval()


def cpu_arm_setup(B):
    """(forward callable, kind, description) of the CPU arm on config 2's weights and one B-sample batch."""
    from oracle import synth, torch_port
    cfg = _oracle_cfg()
    w = synth.make_weights(cfg, seed=42)
    if PRUNED:
        from oracle import prune
        w = prune.one_shot_prune(w, 0.9, 0.444, 1.0)
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    Xi, Xv = synth.make_inputs(cfg, B, seed=0, xv="unit" if XV_UNIT else "int50")
    Xi, Xv = torch.from_numpy(Xi), torch.from_numpy(Xv)
    ref = reference_module(cfg, sd, cuda=False)
    port = lambda: torch_port.forward(cfg, sd, Xi, Xv)        # noqa: E731
    if ref is None:
        return port, "port", "oracle/torch_port.py (the reference's op sequence restated; baseline/_ref is absent)"

    def fwd():
        with torch.no_grad():
            return ref(Xi, Xv)

    # the reference and its restatement agree on this very batch (fp32 rounding only) -- keeps the port honest too
    a, b = fwd(), port()
    assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max()), "baseline/_ref and oracle/torch_port disagree"
    return fwd, "reference", ("the unmodified reference module (baseline/_ref/model/DeepFMs.py, DeepFMs.forward in eval mode under "
                              "torch.no_grad, use_cuda=False)")


def cpu_port_baseline(B, budget_s):
    fwd, kind, what = cpu_arm_setup(B)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd()
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 50):
        t0 = time.perf_counter()
        fwd()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return dict(value=round(B / statistics.median(times), 1), unit=UNIT, cores=cores, kind=kind,
                best=round(B / best, 1),
                sample=f"{len(times)} forwards of the same B={B} batch through {what}, fp32, torch {torch.__version__} CPU, "
                       f"{cores} threads, median")


def reference_cuda(device, B, weights, Xi0, Xv0, iters=20):
    """Supplementary yardstick (SURVEY 8(d) 'optional extra'): the unmodified reference module with use_cuda=True on this same
    B200 (eager aten kernels + cuBLAS SGEMM, model/DeepFMs.py:1012-1022), same weights, same batch, CUDA events around forward."""
    cfg = _oracle_cfg()
    try:
        ref = reference_module(cfg, {k: torch.from_numpy(v) for k, v in weights.items()}, cuda=True)
        if ref is None:
            return {"unavailable": "baseline/_ref is absent"}
        with torch.no_grad():
            out = ref(Xi0, Xv0)
            for _ in range(3):
                ref(Xi0, Xv0)
            torch.cuda.synchronize(device)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                ref(Xi0, Xv0)
            b.record()
            b.synchronize()
        ms = a.elapsed_time(b) / iters
        return dict(value=round(B / (ms * 1e-3), 1), unit=UNIT, ms_per_step=round(ms, 4), iters=iters,
                    what="unmodified reference DeepFMs.forward, use_cuda=True, eager PyTorch on this GPU, fp32 "
                         f"(TF32 off), device-resident inputs, B={B}", logits=out)
    except Exception as ex:
        return {"unavailable": str(ex)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    fwd, kind, what = cpu_arm_setup(B)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    steps = min(args.steps, 60)          # bounded: each step is one full B=4096 forward on the CPU (~0.1-0.4 s)
    for _ in range(min(max(args.warmup, 1), 3)):
        fwd()
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd()
    dt = time.perf_counter() - t0
    v = round(B * steps / dt, 1)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
           "warmup": min(max(args.warmup, 1), 3), "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD_TEXT + f" -- one B={B} batch per step on the host CPU", "batch_per_gpu": B},
           "cpu_baseline": dict(value=v, unit=UNIT, cores=cores, kind=kind,
                                sample=f"{steps} forwards of B={B} through {what}, fp32 torch CPU, {cores} threads"),
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


set_workload("criteo")       # the default for importers (scripts/*.py use make_model / make_batches)

if __name__ == "__main__":
    a = parse()
    set_workload(a.workload)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
