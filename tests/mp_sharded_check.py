"""Run under torchrun (one process per GPU): row-sharded tables (p2p fused gather and the NCCL baseline) must give
bit-identical logits to the unsharded module on the same inputs.  Exit code 0 = pass."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from oracle import closed_form, synth
from oracle.config import PathConfig
from xsdeepfwfm_deprecated_b200.model import DeepFMs
from xsdeepfwfm_deprecated_b200.sharded import ShardedDeepFMs

SIZES = [1] * 13 + [7, 313, 12, 1999, 3, 250, 45, 201, 2, 100003, 77, 5, 640, 9, 33, 4096, 11, 200, 58, 4, 900, 18, 16,
                    129, 89, 250000]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    QR = dict(qr_flag=1, qr_collisions=4, qr_threshold=200)
    # fwlw = False: first order from the fm_1st_embeddings tables (plain and QR), indexed by the original category ids
    for tag, kw, precision, fwlw in (("plain", {}, "fp32", True), ("qr", QR, "fp32", True), ("plain", {}, "bf16x3", True),
                                     ("qr", QR, "bf16x3", True), ("plain1", {}, "bf16x3", False), ("qr1", QR, "bf16x3", False),
                                     ("qr1", QR, "fp32", False)):
        cfg = PathConfig(39, SIZES, use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=fwlw, deep_nodes=64, **kw)
        w = synth.make_weights(cfg, seed=5)
        Xi, Xv = synth.make_inputs(cfg, 777 + 13 * rank, seed=50 + rank)          # ragged, different per rank
        tXi, tXv = torch.from_numpy(Xi).to(dev), torch.from_numpy(Xv).to(dev)
        common = dict(use_fm=False, use_fwfm=True, use_deep=True, use_fwlw=fwlw, deep_nodes=64, use_cuda=True,
                      precision=precision, **kw)
        base = DeepFMs(39, SIZES, **common)
        base.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
        base = base.to(dev).eval()
        with torch.no_grad():
            want = base(tXi, tXv)
        ref = closed_form.forward(cfg, w, Xi, Xv)["logit"]
        assert np.abs(want.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
        for exchange in (("p2p", "p2p_pull", "nccl") if fwlw else ("p2p",)):
            m = ShardedDeepFMs(39, SIZES, exchange=exchange, shard_threshold=200, **common)
            m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
            m = m.to(dev).eval()
            m.shard_()
            nsh = len(m._shards)
            with torch.no_grad():
                got = m(tXi, tXv)
            same = bool(torch.equal(got, want))
            print(f"[rank {rank}/{world}] {tag:5s} {precision:6s} {exchange:8s} sharded_tables={nsh} bit_identical={same}", flush=True)
            ok &= same and nsh >= 4
            m.release()
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
