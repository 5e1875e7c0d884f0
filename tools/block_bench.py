#!/usr/bin/env python
"""Per-kernel CUDA-event timing of ONE Base-L2-shaped block (N=784 tokens, D=192, 8 heads of 24), forward + backward,
for the streamed and the materialised Re-Attention paths side by side.  usage: block_bench.py [B=256] [steps=5]"""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit_unet_b200 as vu
from vit_unet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
vu.set_precision("tf32")
with contextlib.redirect_stdout(io.StringIO()):
    net = vu.HViT_UNet(depth=0, depth_te=1, size_bottleneck=1, preprocessing="conv", im_size=224, patch_size=8,
                       num_channels=3, hidden_dim=32, num_heads=8, attn_drop=0.2, proj_drop=0.2, linear_drop=0)
net.to("cuda").train()
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.rand(B, 3, 224, 224, device="cuda")


def step():
    for p in net.parameters():
        p.grad = None
    vu.l1_loss(net(x), y).backward()


for name, fwd, bwd in (("materialised", False, False), ("streamed fwd", True, False), ("streamed fwd+bwd", True, True)):
    vu.set_streamed(fwd, backward=bwd)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / steps
    t = ops.KernelTimer(); ops.set_kernel_timer(t)
    for _ in range(steps):
        step()
    ops.set_kernel_timer(None)
    by = t.summary(1400.8, 6539.2)["by_kernel"]
    print(f"== {name}: {total:.2f} ms per block step at {B} images ({B / total * 1e3:.0f} img/s for the block alone); peak mem "
          f"{torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GB")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"   {k:48s} {v['ms'] / steps:8.3f} ms  x{v['launches'] // steps:<3d} {v['tflops']:8.1f} TF/s {v['gbs']:8.0f} GB/s")
    torch.cuda.reset_peak_memory_stats()
