"""The oracle (oracle/vit_unet_oracle.py) against the reference's own known answers and against golden
vectors produced by executing the reference's model.py (tests/golden/make_golden.py)."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from make_golden import CONFIGS, fill_state_dict, make_input, pack, run_case
from oracle import vit_unet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _n_params(m):
    return sum(p.numel() for p in m.parameters())


@pytest.mark.parametrize("name,count", [("lite", 3_387_568), ("base", 36_613_036), ("large", 63_043_866)])
def test_readme_param_counts(name, count):
    # README.md:16,34,52 -- reproduced only by 3x3 q/k/v convs + one shared LN + PE conv + fine table
    assert _n_params(O.get_vit_unet(name, variant="readme")) == count


@pytest.mark.parametrize("name,count", [("lite", 5_193_820), ("base", 39_623_512), ("large", 69_064_902)])
def test_head_param_counts(name, count):
    # SURVEY.md F2 [probe]: V-HEAD (separate LN1/LN2) parameter counts
    assert _n_params(O.get_vit_unet(name, variant="head")) == count


def test_head_1ch_param_count():
    assert _n_params(O.get_vit_unet("base", variant="head", num_channels=1)) == 6_228_718


def test_notebook_shapes():
    # ViT_UNet.ipynb c8 / c18 / c32 / c37 / c47: 512^2 image, p=32
    x = torch.arange(3 * 512 * 512, dtype=torch.float32).reshape(1, 3, 512, 512)
    t = O.patchify(x, 32)
    assert t.shape == (1, 256, 3072)
    d = O.resample(t, 3, 0.5)
    assert d.shape == (1, 1024, 768)
    assert O.resample(d, 3, 0.5).shape == (1, 4096, 192)
    assert O.resample(t, 3, 2).shape == (1, 64, 12288)
    assert torch.equal(O.unpatchify(t, 3), x)
    assert torch.equal(O.resample(d, 3, 2), t)
    assert torch.equal(d, O.patchify(x, 16))
    # token r*G+c, feature ch*p*p+i*p+j <-> pixel (ch, r*p+i, c*p+j)   (SURVEY A1 [probe 1a])
    assert t[0, 11 * 16 + 3, 2 * 1024 + 5 * 32 + 7] == x[0, 2, 11 * 32 + 5, 3 * 32 + 7]


def test_bad_preset_and_asserts():
    with pytest.raises(ValueError):
        O.get_vit_unet("huge")
    with pytest.raises(AssertionError):
        O.HViT_UNet(3, 1, 1, "conv", 224, 16, 3, 64, 4, 0., 0., 0)      # final patch 2 < 4


def test_state_dict_keys_match_reference_layout():
    m = O.get_vit_unet("lite", variant="head")
    keys = set(m.state_dict().keys())
    for k in ["PE.position_embedding.weight", "Encoders.0.ReAttn.reatten_matrix.weight",
              "Encoders.0.ReAttn.var_norm.running_mean", "Encoders.0.ReAttn.var_norm.num_batches_tracked",
              "Encoders.0.ReAttn.qconv2d.weight", "Encoders.0.ReAttn.proj.bias", "Encoders.0.LN1.weight",
              "Encoders.0.LN2.bias", "Encoders.0.FeedForward.net.0.weight", "Encoders.0.FeedForward.net.3.bias",
              "BottleNeck.1.ReAttn.kconv2d.weight", "Decoders.1.LN1.weight",
              "SkipConnections.1.vconv2d.weight", "SkipConnections.0.proj.weight", "conv2d.weight", "conv2d.bias"]:
        assert k in keys, k
    assert m.state_dict()["Encoders.0.LN1.weight"].shape == (196, 768)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_matches_reference_golden(name):
    variant, kw, B = CONFIGS[name]
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = O.HViT_UNet(**kw)
    assert _n_params(model) == int(gold["n_params"])
    model.load_state_dict(fill_state_dict(model.state_dict()))
    x, y = make_input(B, kw["num_channels"], kw["im_size"])
    got = pack(run_case(model, x, y, train=True), full=name.startswith("tiny"))
    if not name.startswith(("lite", "base")):          # MSE-loss vectors (tags mev / mtr): the small configs bound the CPU time
        model.load_state_dict(fill_state_dict(model.state_dict()))
        got.update({k: v for k, v in pack(run_case(model, x, y, train=True, loss_kind="mse"), False).items()
                    if k.startswith(("mev_", "mtr_"))})
    skip = lambda k: "_cond:" in k or k.startswith("r64:") or (k.startswith(("mev_", "mtr_")) and k not in got)
    assert set(got) | {"n_params"} == {k for k in gold.files if not skip(k)}
    for k in gold.files:
        if k == "n_params" or skip(k):      # r64: the reference evaluated in fp64
            continue
        g, o = gold[k], got[k]
        scale = max(np.abs(g).max(), 1e-30)
        # same ATen kernels, different op grouping (batched conv, reshape-based patchify): fp32 round-off only
        assert np.abs(o - g).max() <= 2e-5 * scale + 1e-7, (k, np.abs(o - g).max(), scale)


def test_dice_matches_readme_formula():
    g = torch.Generator().manual_seed(3)
    a, t = torch.rand(2, 1, 8, 8, generator=g), (torch.rand(2, 1, 8, 8, generator=g) > 0.5).float()
    exp = 1 - (2 * (a * t).sum() + 1) / (a.sum() + t.sum() + 1)
    assert torch.allclose(O.dice_loss(a, t), exp)
