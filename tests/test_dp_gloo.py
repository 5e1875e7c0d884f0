"""World-size-2 gloo test (CPU) of the data-parallel host logic: suffix bucketing of the flat gradient buffer
and averaged all-reduce.  The kernels are not involved (they need a GPU); a fake backward fills the groups in
the order the engine's schedule finishes them."""
import contextlib
import io
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vit_unet_b200 as vu
        from vit_unet_b200.dp import DataParallel, _group_of
        with contextlib.redirect_stdout(io.StringIO()):
            net = vu.HViT_UNet(depth=2, depth_te=2, size_bottleneck=1, preprocessing="conv", im_size=32,
                               patch_size=16, num_channels=3, hidden_dim=32, num_heads=4, attn_drop=0.,
                               proj_drop=0., linear_drop=0)
        torch.manual_seed(100 + rank)                    # ranks start with DIFFERENT weights ...
        for p in net.parameters():
            p.data.normal_()
        dpm = DataParallel(net, bucket_mb=0.5)           # ... and the wrapper broadcasts rank 0's
        ref0 = [p.detach().clone() for p in net.parameters()]
        gathered = [torch.empty_like(ref0[0]) for _ in range(world)]
        dist.all_gather(gathered, ref0[0])
        assert all(torch.equal(g, gathered[0]) for g in gathered)

        b = dpm.bucketer
        flat = torch.zeros(net._flat_numel)
        b.begin(flat)
        # fake backward: groups finish in reverse forward order, each rank writes rank-dependent values
        order, seen = [], set()
        for n in net._param_names:
            g = _group_of(n)
            if g not in seen:
                seen.add(g); order.append(g)
        starts = b.group_starts
        ends = {g: (starts[order[i + 1]] if i + 1 < len(order) else net._flat_numel) for i, g in enumerate(order)}
        for g in reversed(order):
            flat[starts[g]:ends[g]] = float(rank + 1) * (1 + order.index(g))
            b.on_ready(g)
        b.finish()
        exp = torch.zeros_like(flat)
        for g in order:
            exp[starts[g]:ends[g]] = (sum(range(1, world + 1)) / world) * (1 + order.index(g))
        assert torch.allclose(flat, exp), "averaged gradients differ"
        # buckets: contiguous, descending, cover the whole buffer exactly once, more than one of them
        spans = b.launched
        assert len(spans) > 1 and spans[0][1] == net._flat_numel and spans[-1][0] == 0
        assert all(spans[i][0] == spans[i + 1][1] for i in range(len(spans) - 1))
        assert order[0] == "PE." and order[-1] == "conv2d."
        # the projection weight of every attention opens its own notification point (early bucket for the largest tensor)
        pd = dict(zip(net._param_names, net._flat_offsets))
        assert b.group_starts["Encoders.0.ReAttn.proj."] == pd["Encoders.0.ReAttn.proj.weight"]
        assert b.group_starts["SkipConnections.0.proj."] == pd["SkipConnections.0.proj.weight"]
        b.begin(flat)
        b.bucket_numel = 1
        b.on_ready("Encoders.0.ReAttn.proj.")
        assert b.launched == [(pd["Encoders.0.ReAttn.proj.weight"], net._flat_numel)]
        b.finish()
        # BatchNorm running statistics drift apart per rank during training; checkpointing broadcasts rank 0's
        bufs = dict(net.named_buffers())
        name = next(n for n in bufs if n.endswith("var_norm.running_mean"))
        bufs[name].fill_(float(rank + 7))
        if rank == 0:                                    # the usual rank-0-only checkpoint: no collective, no hang
            sd0 = dpm.state_dict()
            assert torch.all(sd0[name] == 7.0)
        dpm.sync_buffers(0)                              # explicit collective on every rank
        sd = dpm.state_dict()                            # the reference's key layout
        assert name in sd and not any(k.startswith("module.") for k in sd)
        assert torch.all(sd[name] == 7.0), sd[name]
        # src selects the broadcasting rank (a rank inside the process group; the default group here)
        bufs[name].fill_(float(rank + 20))
        dpm.sync_buffers(1)
        assert torch.all(bufs[name] == 21.0)
        # a model without an output conv has no 'conv2d.' group: Engine.backward still notifies it
        with contextlib.redirect_stdout(io.StringIO()):
            net2 = vu.ViT_UNet(depth=1, depth_te=1, size_bottleneck=1, preprocessing="none", num_patches=4, patch_size=8,
                               num_channels=3, hidden_dim=16, num_heads=2, attn_drop=0., proj_drop=0., linear_drop=0)
        dp2 = DataParallel(net2, bucket_mb=0.001)
        flat2 = torch.full((net2._flat_numel,), float(rank + 1))
        dp2.bucketer.begin(flat2)
        assert "conv2d." not in dp2.bucketer.group_starts
        dp2.bucketer.on_ready("conv2d.")                 # must be a no-op, not a KeyError
        for st in reversed(net2.engine.sched):
            if st[0] in ("block", "skip"):
                dp2.bucketer.on_ready(st[1])
        dp2.bucketer.on_ready("PE.")
        dp2.bucketer.finish()
        assert torch.allclose(flat2, torch.full_like(flat2, 1.5))
        q.put((rank, "ok"))
    except Exception as e:       # noqa: BLE001
        q.put((rank, f"fail: {e!r}"))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
