"""N3 on the device: converted checkpoints loaded into the CUDA modules reproduce the outputs, and the
best-checkpoint.bin save / reload flow of run_denoising.py:88,100 restores model and optimizer state."""
import contextlib
import io

import pytest
import torch

pytestmark = pytest.mark.gpu

from make_golden import fill_state_dict, make_input       # noqa: E402

COMMON = dict(depth=2, depth_te=1, size_bottleneck=1, preprocessing="conv", patch_size=16, num_channels=3,
              hidden_dim=32, num_heads=4, attn_drop=0., proj_drop=0., linear_drop=0)
GEOM = dict(num_channels=3, im_size=32, patch_size=16, depth=2)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_converted_checkpoint_runs_on_the_cuda_modules():
    import vit_unet_b200 as vu
    from vit_unet_b200 import checkpoint as ck
    head = _quiet(vu.HViT_UNet, im_size=32, **COMMON)
    readme = _quiet(vu.ViT_UNet, num_patches=4, **COMMON)
    sd = fill_state_dict(head.state_dict())
    for k in list(sd):
        if ".LN2." in k:
            sd[k] = sd[k.replace(".LN2.", ".LN1.")].clone()
    head.load_state_dict(sd)
    readme.load_state_dict(ck.head_to_readme(sd, **GEOM))
    head.to("cuda").eval(); readme.to("cuda").eval()
    x, _ = make_input(2, 3, 32)
    with torch.no_grad():
        a, b = head(x.cuda()), readme(x.cuda())
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    back = ck.readme_to_head({k: v.cpu() for k, v in readme.state_dict().items()}, **GEOM)
    assert all(torch.equal(back[k], sd[k]) for k in sd)


def test_best_checkpoint_save_and_reload(tmp_path):
    import vit_unet_b200 as vu
    from vit_unet_b200 import checkpoint as ck
    from vit_unet_b200.dp import DataParallel
    net = _quiet(vu.HViT_UNet, im_size=32, **COMMON)
    net.load_state_dict(fill_state_dict(net.state_dict()))
    net.to("cuda").train()
    model = DataParallel(net)                                   # world size 1: same code path as the multi-GPU run
    opt = vu.FusedAdamW(net.parameters(), lr=1e-3).flatten(net)
    x, y = make_input(2, 3, 32)
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        vu.mse_loss(model(x.cuda()), y.cuda()).backward()
        opt.step()
    path = str(tmp_path / "best-checkpoint.bin")
    ck.save_checkpoint(path, model, optimizer=opt, best_summary_loss=0.5, epoch=3)
    net.eval()
    with torch.no_grad():
        ref = net(x.cuda())
    net2 = _quiet(vu.HViT_UNet, im_size=32, **COMMON).to("cuda")
    opt2 = vu.FusedAdamW(net2.parameters(), lr=1e-3).flatten(net2)
    blob = ck.load_checkpoint(path, net2, optimizer=opt2)
    assert blob["epoch"] == 3 and blob["best_summary_loss"] == 0.5
    net2.eval()
    with torch.no_grad():
        assert torch.allclose(net2(x.cuda()), ref, rtol=1e-6, atol=1e-7)   # weights AND BatchNorm running statistics restored
    assert int(opt2._flat["step"]) == 2
    assert torch.equal(opt2._flat["m"], opt._flat["m"]) and torch.equal(opt2._flat["v"], opt._flat["v"])
    # the README-variant module loads the same file through the layout conversion (LN1 == LN2 is required: refused here)
    readme = _quiet(vu.ViT_UNet, num_patches=4, **COMMON).to("cuda")
    with pytest.raises(ValueError):
        ck.load_checkpoint(path, readme, geometry=GEOM)
