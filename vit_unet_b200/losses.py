"""Fused scalar losses on the CUDA path: one reduction kernel forward, one elementwise kernel backward.

* ``mse_loss``  == torch.nn.MSELoss()        (run_denoising.py:80)
* ``dice_loss`` == the README snippet        (README.md:91-101: smooth=1, whole batch flattened, raw outputs)
* ``l1_loss``   == torch.nn.L1Loss()         (named by the benchmark configs)
"""
from __future__ import annotations

import torch

from . import ops


def _world(group):
    import torch.distributed as dist
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, kind, group=None, global_reduce=False):
        pred = pred.contiguous().float()
        target = target.contiguous().float()
        if pred.shape != target.shape:
            raise ValueError(f"loss: shape mismatch {tuple(pred.shape)} vs {tuple(target.shape)}")
        sums = torch.empty(4, dtype=torch.float64, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        ops.loss_fwd(kind, pred, target, sums, loss)
        ctx.kind, ctx.gmul = kind, 1.0
        world = _world(group) if global_reduce else 1
        if world > 1:
            # data parallel, whole-batch semantics: the ranks add their sums (3 doubles), every rank gets the GLOBAL loss;
            # backward uses the global sums, and because the DataParallel wrapper AVERAGES gradients over ranks the local
            # gradient is pre-multiplied by the world size so that the average is the gradient of the global loss
            import torch.distributed as dist
            dist.all_reduce(sums, group=group)
            ops.loss_finalize(kind, pred.numel() * world, sums, loss)
            ctx.gmul = float(world)
        ctx.save_for_backward(pred, target, sums)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, target, sums = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        gs = g.contiguous().float().reshape(1)
        if ctx.gmul != 1.0:
            gs = gs * ctx.gmul
        ops.loss_bwd(ctx.kind, pred, target, sums, gs, dpred)
        return dpred, None, None, None, None


def l1_loss(pred, target):
    return _LossFn.apply(pred, target, "l1")


def mse_loss(pred, target):
    return _LossFn.apply(pred, target, "mse")


def dice_loss(input, target, group=None, global_batch=True):
    """README.md:91-101 soft-Dice: 1 - (2 sum(xy) + 1) / (sum x + sum y + 1) over the WHOLE batch (a ratio, not a mean).
    Under data parallelism (torch.distributed initialised, world > 1) `global_batch=True` keeps that definition: the
    three sums are all-reduced over `group` (24 bytes), every rank returns the global Dice, and gradients are scaled
    so that the DataParallel gradient average equals the gradient of the global loss.  `global_batch=False` gives the
    per-rank Dice (mean over ranks of local ratios after gradient averaging)."""
    return _LossFn.apply(input, target, "dice", group, global_batch)


class L1Loss(torch.nn.Module):
    def forward(self, pred, target):
        return l1_loss(pred, target)


class MSELoss(torch.nn.Module):
    def forward(self, pred, target):
        return mse_loss(pred, target)


class DiceLoss(torch.nn.Module):
    def forward(self, pred, target):
        return dice_loss(pred, target)
