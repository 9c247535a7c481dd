"""Debug tool (torchrun, >= 2 GPUs): do the row-pull kernel and the fused kernel overlap on the device?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from oracle import synth
from xsdeepfwfm_deprecated_b200 import _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
B, nb = 4096, 8
m = bench.make_model(dev, "bf16x3", synth.CRITEO_PAPER, world)
Xi, Xv = bench.make_batches(dev, synth.CRITEO_PAPER, B, nb, seed=rank)
lanes = [m.pull_lane(B, k).prepare(m, "bf16x3") for k in range(2)]
out = torch.empty(nb, B, device=dev)
s_f, s_p = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
s_p0 = torch.cuda.Stream(dev)

def pull(l, j, st): lanes[l].enqueue_pull(lib, Xi[j].data_ptr(), 26, 1, st.cuda_stream)
def fwd(l, j, st): lanes[l].enqueue_forward(lib, Xv[j].data_ptr(), 13, 1, out[j].data_ptr(), None, st.cuda_stream)

def timed(fn, n=50):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    a.record(main)
    for i in range(n):
        fn(i)
    s_f.synchronize(); s_p.synchronize(); s_p0.synchronize()
    b.record(main); b.synchronize()
    return a.elapsed_time(b) * 1000 / n

for l in range(2):
    pull(l, 0, s_f); fwd(l, 0, s_f)
torch.cuda.synchronize()
t_pull = timed(lambda i: pull(0, i % nb, s_f))
t_fwd = timed(lambda i: fwd(0, i % nb, s_f))
t_both_hi = timed(lambda i: (fwd(0, i % nb, s_f), pull(1, i % nb, s_p)))
t_both_eq = timed(lambda i: (fwd(0, i % nb, s_f), pull(1, i % nb, s_p0)))
if rank == 0:
    print(f"pull alone {t_pull:.1f} us, fused alone {t_fwd:.1f} us (host-launched, wall clock incl. launch gaps)")
    print(f"fused (stream A) || pull (high-priority stream): {t_both_hi:.1f} us per pair; equal priority: {t_both_eq:.1f} us per pair")
m.release(); dist.destroy_process_group()
