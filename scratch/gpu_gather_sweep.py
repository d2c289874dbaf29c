"""torchrun script: resample_batch_gathered (resample + exchange pipelined over sub-batches) against the sub-batch count"""
import os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import soxr
from soxr import distributed as sdist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.randn((1024, 48000, 2), device=dev)
def timed(f, n=5, warm=2):
    for _ in range(warm): f()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
y = soxr.resample_batch(x, 48000, 44100, 'VHQ')
full = torch.empty((world * 1024,) + tuple(y.shape[1:]), device=dev)
t_c = timed(lambda: soxr.resample_batch(x, 48000, 44100, 'VHQ', out=y))
t_g = timed(lambda: dist.all_gather_into_tensor(full, y))
if rank == 0: print("world %d: compute %.3f ms, all_gather_into_tensor of the whole shard %.3f ms" % (world, t_c, t_g), flush=True)
for method in ('peer', 'nccl'):
    for sb in (1, 2, 4, 8, 16):
        try:
            t = timed(lambda: sdist.resample_batch_gathered(x, 48000, 44100, 'VHQ', sub_batches=sb, method=method))
        except Exception as e:
            if rank == 0: print(method, "failed:", repr(e)[:300], flush=True)
            break
        if rank == 0: print("%s sub_batches %2d: %.3f ms per step" % (method, sb, t), flush=True)
# correctness of the pushed result
full = sdist.resample_batch_gathered(x, 48000, 44100, 'VHQ', sub_batches=4, method='auto')
ref = torch.empty_like(full)
dist.all_gather_into_tensor(ref, soxr.resample_batch(x, 48000, 44100, 'VHQ'))
torch.cuda.synchronize()
if rank == 0: print("gathered result equals all_gather_into_tensor of the shards:", bool(torch.equal(full, ref)), flush=True)
dist.destroy_process_group()
