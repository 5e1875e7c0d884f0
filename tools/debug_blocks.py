"""Debug helper (GPU box): per-block comparison of the CUDA engine against the oracle in train mode."""
import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
from make_golden import CONFIGS, fill_state_dict, make_input
from oracle import vit_unet_oracle as O
import vit_unet_b200 as vu

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_head_te2"
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
variant, kw, B = CONFIGS[name]
with contextlib.redirect_stdout(io.StringIO()):
    ref, net = O.HViT_UNet(**kw), vu.HViT_UNet(**kw)
sd = fill_state_dict(ref.state_dict())
ref.load_state_dict(sd); net.load_state_dict(sd); net.to("cuda")
x, y = make_input(B, kw["num_channels"], kw["im_size"])
ref.train(mode == "train"); net.train(mode == "train")
acts = {}
def hook(nm):
    def f(m, i, o): acts[nm] = o.detach()
    return f
for nm, m in ref.named_modules():
    if nm.count(".") == 1 and nm.split(".")[0] in ("Encoders", "BottleNeck", "Decoders", "SkipConnections"):
        m.register_forward_hook(hook(nm + "."))
    if nm.endswith("ReAttn"):
        m.register_forward_hook(hook(nm + ".attnout"))
out_ref = ref(x)
P = {n: p.detach() for n, p in net.named_parameters()}; P.update(dict(net.named_buffers()))
out, saved = net.engine.forward(P, x.cuda(), train=(mode == "train"), save=True, seed=0)
def rel(a, b):
    a, b = a.cpu().double(), b.cpu().double()
    return ((a - b).abs().max() / b.abs().max()).item()
steps = [s for s in net.engine.sched]
for st, sv in zip(steps, saved["steps"]):
    if st[0] == "block":
        pre = st[1]
        # y1 = attn_out + x  -> compare LN1 input indirectly via x1; and final via next
        print(pre, "x1 vs ?", end=" ")
        print("y2-stat", sv["st2"].cpu().tolist()[:1], end=" ")
        a = acts[pre]
        # block output is not saved directly; recompute from y2/st2 not needed: compare attn out
        att = acts[pre + "ReAttn.attnout"]
        y1 = sv["y1"]; xin = sv["attn"]["xq"]
        print("attn_out rel", rel(y1 - xin, att))
    elif st[0] == "skip":
        pre = st[1]
        print(pre, "skip out: (checked through final)")
print("final out rel", rel(out, out_ref))
