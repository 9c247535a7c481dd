"""Debug tool: the fused forward reading Xi/Xv straight from pinned host memory (UVA) and writing prob to it,
against the staged-copy streamed API.  Usage: python scripts/zerocopy_probe.py [precision] [nstreams]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
nstreams = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, nh = 4096, 32
sizes = synth.CRITEO_PAPER
P = _lib.PRECISIONS[prec]
dXi, dXv = bench.make_batches(dev, sizes, B, nh, seed=0)
for idt in ("int64", "int32"):
    m = bench.make_model(dev, prec, sizes, index_dtype=idt)
    plan = m._get_plan(); plan.ensure_image(m, prec)
    hXi = dXi[:, :, :, 0].cpu().to(torch.int32 if idt == "int32" else torch.int64).pin_memory()
    hXv = dXv.cpu().pin_memory()
    hout = torch.empty(nh, B).pin_memory()
    hout2 = torch.empty(nh, B).pin_memory()
    ws = torch.zeros(lib.dfw_forward_host_stream_workspace_bytes(plan.model_ref, B, P) + 4096, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    dout = torch.empty(nh, B, device=dev)

    def staged():
        _lib.check(lib.dfw_forward_host_stream(plan.model_ref, hXi.data_ptr(), hXv.data_ptr(), nh * B, B, P, ws.data_ptr(),
                                               ws.numel(), None, hout.data_ptr(), st), "staged")

    pxi = [hXi[j].data_ptr() for j in range(nh)]
    pxv = [hXv[j].data_ptr() for j in range(nh)]
    pout = {True: [hout2[j].data_ptr() for j in range(nh)], False: [dout[j].data_ptr() for j in range(nh)]}
    sps = [s.cuda_stream for s in streams]
    fwd, mref = lib.dfw_forward_fused, plan.model_ref

    def zero(out_host=True):
        po = pout[out_host]
        for j in range(nh):
            rc = fwd(mref, pxi[j], 26, 1, pxv[j], 13, 1, B, P, None, po[j], None, sps[j % nstreams])
            if rc:
                _lib.check(rc, "zero")
        torch.cuda.synchronize()

    for name, fn in (("staged copies", staged), ("zero-copy in+out", zero), ("zero-copy in, device out", lambda: zero(False))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / (10 * nh)
        print(f"{prec} {idt} {name}: {t * 1e6:.1f} us/step -> {B / t / 1e6:.1f} M samples/s", flush=True)
    print("outputs equal:", torch.equal(hout, hout2), flush=True)
