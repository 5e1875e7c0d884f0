"""N3: state_dict conversion between the README and HEAD variants is function-preserving (checked with the CPU
oracle, eval mode) and refuses inexact conversions."""
import contextlib
import io

import pytest
import torch

from make_golden import fill_state_dict, make_input
from oracle import vit_unet_oracle as O
from vit_unet_b200 import checkpoint as ck

GEOM = dict(num_channels=3, im_size=32, patch_size=16, depth=2)


def _models():
    common = dict(depth=2, depth_te=1, size_bottleneck=1, preprocessing="conv", patch_size=16, num_channels=3,
                  hidden_dim=32, num_heads=4, attn_drop=0., proj_drop=0., linear_drop=0)
    with contextlib.redirect_stdout(io.StringIO()):
        head = O.HViT_UNet(im_size=32, **common)
        readme = O.ViT_UNet(num_patches=4, **common)
    return head, readme


def test_head_to_readme_and_back_preserves_the_function():
    head, readme = _models()
    sd = fill_state_dict(head.state_dict())
    for k in list(sd):                      # make LN1 == LN2 so the conversion is exact
        if ".LN2." in k:
            sd[k] = sd[k.replace(".LN2.", ".LN1.")].clone()
    head.load_state_dict(sd)
    readme.load_state_dict(ck.head_to_readme(sd, **GEOM))
    x, _ = make_input(2, 3, 32)
    head.eval(); readme.eval()
    with torch.no_grad():
        a, b = head(x), readme(x)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    back = ck.readme_to_head(readme.state_dict(), **GEOM)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


def test_inexact_conversions_are_refused():
    head, readme = _models()
    with pytest.raises(ValueError):          # LN1 != LN2
        ck.head_to_readme(fill_state_dict(head.state_dict()), **GEOM)
    with pytest.raises(ValueError):          # non-identity PE conv
        ck.readme_to_head(fill_state_dict(readme.state_dict()), **GEOM)
