"""One full Base training step (BASELINE configs[2]: zero_grad + forward + L1 loss + backward, preset dropout) at B images
between cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures of every kernel of the benchmarked step."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit_unet_b200 as vu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vu.set_precision(sys.argv[2] if len(sys.argv) > 2 else "bf16")      # the bench default
with contextlib.redirect_stdout(io.StringIO()):
    net = vu.get_vit_unet("base")
net.to("cuda").train()
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.rand(B, 3, 224, 224, device="cuda")
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for p in net.parameters(): p.grad = None
    vu.l1_loss(net(x), y).backward()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
