import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import torch
from make_golden import CONFIGS, fill_state_dict, make_input
from oracle import vit_unet_oracle as O
import vit_unet_b200 as vu

name = sys.argv[1] if len(sys.argv) > 1 else "tiny_head_te2"
do_eval_first = (sys.argv[2] if len(sys.argv) > 2 else "1") == "1"
variant, kw, B = CONFIGS[name]
with contextlib.redirect_stdout(io.StringIO()):
    ref, net = O.HViT_UNet(**kw), vu.HViT_UNet(**kw)
sd = fill_state_dict(ref.state_dict())
ref.load_state_dict(sd); net.load_state_dict(sd); net.to("cuda")
x, y = make_input(B, kw["num_channels"], kw["im_size"])
def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max()).item()
if do_eval_first:
    ref.eval(); net.eval()
    with torch.no_grad():
        print("eval rel", rel(net(x.cuda()), ref(x)))
ref.train(); net.train()
xr = x.clone().requires_grad_(True); xn = x.clone().cuda().requires_grad_(True)
o_r = ref(xr); o_n = net(xn)
print("train out rel", rel(o_n, o_r))
lr = torch.nn.functional.l1_loss(o_r, y)
ln = vu.l1_loss(o_n, y.cuda())
lt = torch.nn.functional.l1_loss(o_n.detach().cpu(), y)
print("loss ref", lr.item(), "ours", ln.item(), "torch-on-ours", lt.item())
lr.backward(); ln.backward()
print("loss after bwd: ours", ln.item())
print("dx rel", rel(xn.grad, xr.grad))
gr = dict(ref.named_parameters())
rows = sorted(((rel(p.grad, gr[n].grad), n) for n, p in net.named_parameters() if not n.endswith("reatten_matrix.bias")), reverse=True)
for r, n in rows[:8]:
    print(f"  {r:.3e} {n}")
