#!/usr/bin/env python
"""Timeline evidence for the overlapped gradient all-reduce (no nsys in the image: torch.profiler / CUPTI instead).
Run under torchrun with N >= 2 ranks:  python -m torch.distributed.run --nproc-per-node N tools/dp_timeline.py [batch]
Rank 0 profiles ONE training step and prints, for every NCCL kernel, when it started relative to the first kernel of the
step, how long it ran, and how much of it lay behind the last compute kernel (the un-overlapped tail)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vit_unet_b200 as vu
from vit_unet_b200.dp import DataParallel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
vu.set_precision("tf32")
with contextlib.redirect_stdout(io.StringIO()):
    net = vu.get_vit_unet("base")
net.to("cuda").train()
model = DataParallel(net)
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.rand(B, 3, 224, 224, device="cuda")


def step():
    for p in net.parameters():
        p.grad = None
    vu.l1_loss(model(x), y).backward()


for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.cuda_time_total is not None]
    ks = [(e.time_range.start, e.time_range.end, e.name) for e in ev if e.time_range.end > e.time_range.start]
    ks.sort()
    t0 = ks[0][0]
    nccl = [k for k in ks if "nccl" in k[2].lower()]
    comp = [k for k in ks if "nccl" not in k[2].lower() and "memcpy" not in k[2].lower() and "memset" not in k[2].lower()]
    last_comp = max(k[1] for k in comp)
    print(f"step: {len(comp)} compute kernels, {(last_comp - t0) / 1e3:.2f} ms from first to last compute kernel; world {dist.get_world_size()}, {B} img/GPU")
    print(f"buckets launched: {model.bucketer.launched} (flat gradient buffer of {net._flat_numel} floats)")
    for s, e, n in nccl:
        tail = max(0.0, e - max(s, last_comp))
        print(f"  NCCL {n[:48]:48s} start +{(s - t0) / 1e3:8.2f} ms  duration {(e - s) / 1e3:6.2f} ms  behind the last compute kernel: {tail / 1e3:5.2f} ms")
    end = max(k[1] for k in ks)
    print(f"step end (all kernels): +{(end - t0) / 1e3:.2f} ms; un-overlapped communication tail: {(end - last_comp) / 1e3:.2f} ms")
dist.destroy_process_group()
