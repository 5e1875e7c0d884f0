"""Lite / Base inference (eval forward, no grad) at B images between cudaProfilerStart/Stop, for ncu captures of the
streamed forward kernel.  usage: profile_infer.py [lite|base] [B]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vit_unet_b200 as vu

preset = sys.argv[1] if len(sys.argv) > 1 else "lite"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
vu.set_precision("tf32")
with contextlib.redirect_stdout(io.StringIO()):
    net = vu.get_vit_unet(preset)
net.to("cuda").eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
with torch.no_grad():
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize(); torch.cuda.profiler.start()
        net(x)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
