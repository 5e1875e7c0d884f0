"""One Base-L2-shaped block (N=784 tokens, D=192, 8 heads, hd=24: where 92% of the attention-map bytes live),
forward + backward at B images, for `ncu --profile-from-start off` captures of every kernel of the block."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit_unet_b200 as vu

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
vu.set_precision(prec)
with contextlib.redirect_stdout(io.StringIO()):
    net = vu.HViT_UNet(depth=0, depth_te=1, size_bottleneck=1, preprocessing="conv", im_size=224, patch_size=8,
                       num_channels=3, hidden_dim=32, num_heads=8, attn_drop=0.2, proj_drop=0.2, linear_drop=0)
net.to("cuda").train()
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.rand(B, 3, 224, 224, device="cuda")
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for p in net.parameters(): p.grad = None
    vu.l1_loss(net(x), y).backward()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
