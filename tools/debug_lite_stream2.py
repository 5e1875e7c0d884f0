"""Streamed vs materialised training path on small 4-head models, both judged against the FP32-exact path of the same
model (is a streamed-vs-materialised gap a kernel error or the conditioning of the configuration?)."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import torch
import vit_unet_b200 as vu
from make_golden import fill_state_dict, make_input

for im, p, C, heads, B, depth in ((224, 8, 3, 4, 1, 1), (112, 8, 3, 4, 2, 1), (224, 16, 3, 4, 1, 2), (224, 16, 3, 4, 2, 2)):
    kw = dict(depth=depth, depth_te=1, size_bottleneck=1, preprocessing="conv", im_size=im, patch_size=p, num_channels=C,
              hidden_dim=16, num_heads=heads, attn_drop=0.0, proj_drop=0.0, linear_drop=0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = vu.HViT_UNet(**kw)
    net.load_state_dict(fill_state_dict(net.state_dict()))
    net.to("cuda")
    x, y = make_input(B, C, im)
    x, y = x.cuda().requires_grad_(True), y.cuda()
    for mode in (False, True):
        res = {}
        for tag, prec, st in (("fp32", "fp32", False), ("mat", "tf32", False), ("str", "tf32", True)):
            vu.set_precision(prec)
            vu.set_streamed(st)
            net.train(mode); net.zero_grad(); x.grad = None
            net.load_state_dict(fill_state_dict(net.state_dict()))
            vu.mse_loss(net(x), y).backward()
            res[tag] = {n: p_.grad.detach().clone() for n, p_ in net.named_parameters()}
            res[tag]["dx"] = x.grad.detach().clone()
        ref = res["fp32"]
        for tag in ("mat", "str"):
            b = res[tag]
            worst = sorted(((float((b[n] - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-30)), n) for n in ref), reverse=True)[:3]
            print(f"depth={depth} im={im} p={p} B={B} train={mode} {tag} vs fp32: worst", [(round(e, 4), n) for e, n in worst], flush=True)
