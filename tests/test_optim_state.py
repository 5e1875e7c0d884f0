"""FusedAdamW checkpoint/resume (CPU: no kernel is launched): after flatten() the per-parameter state entries are
views of the flat moment buffers and share one step counter, so optimizer.state_dict() / load_state_dict() round-trip
the single-launch path (the reference fitter stores optimizer_state_dict)."""
import contextlib
import io

import torch

import vit_unet_b200 as vu


def _net():
    with contextlib.redirect_stdout(io.StringIO()):
        return vu.HViT_UNet(depth=1, depth_te=1, size_bottleneck=1, preprocessing="conv", im_size=16, patch_size=8,
                            num_channels=3, hidden_dim=16, num_heads=2, attn_drop=0., proj_drop=0., linear_drop=0)


def test_flat_state_round_trips_through_state_dict():
    net = _net()
    opt = vu.FusedAdamW(net.parameters(), lr=1e-3).flatten(net)
    f = opt._flat
    g = torch.Generator().manual_seed(3)
    f["m"].copy_(torch.rand(f["m"].shape, generator=g)); f["v"].copy_(torch.rand(f["v"].shape, generator=g))
    f["step"].fill_(17)
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(net.parameters()))            # nothing is hidden outside self.state
    assert all(int(s["step"]) == 17 for s in sd["state"].values())
    net2 = _net()
    opt2 = vu.FusedAdamW(net2.parameters(), lr=1e-3).flatten(net2)
    opt2.load_state_dict(sd)
    assert int(opt2._flat["step"]) == 17
    pd, pd2 = dict(net.named_parameters()), dict(net2.named_parameters())
    for name, off in zip(net2._param_names, net2._flat_offsets):
        n = pd2[name].numel()
        assert torch.equal(opt2._flat["m"][off:off + n], f["m"][off:off + n]), name
        assert torch.equal(opt2._flat["v"][off:off + n], f["v"][off:off + n]), name
        st = opt2.state[pd2[name]]                                    # still views of the flat buffers after loading
        assert st["m"].data_ptr() == opt2._flat["m"].data_ptr() + 4 * off and st["step"] is opt2._flat["step"]
    assert all(p.data.data_ptr() == opt._flat["p"].data_ptr() + 4 * o
               for (n_, o) in zip(net._param_names, net._flat_offsets) for p in [pd[n_]])
