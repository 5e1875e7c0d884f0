"""Checkpoint interop between the two reference variants.

V-HEAD (``HViT_UNet``: LN1/LN2 per block, no PatchEncoder conv, position table indexed at the coarsest patch size,
model.py:80-82,193-196) and V-README (``ViT_UNet``: one shared LN per block, PatchEncoder conv, table indexed at the
finest patch size, ViT_UNet.ipynb c16/c27) describe the same network family; this module converts state_dicts
between them whenever the conversion is exact, and refuses otherwise.  Checkpoint tooling, not hot path: plain
torch on whatever device the tensors live on.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def _repatch_table(table: torch.Tensor, C: int, S: int, p_from: int, p_to: int) -> torch.Tensor:
    """(N_from, C*p_from^2) position table -> the same per-pixel offsets laid out for patch size p_to."""
    g = S // p_from
    img = table.reshape(g, g, C, p_from, p_from).permute(2, 0, 3, 1, 4).reshape(C, S, S)
    g2 = S // p_to
    return img.reshape(C, g2, p_to, g2, p_to).permute(1, 3, 0, 2, 4).reshape(g2 * g2, C * p_to * p_to).contiguous()


def readme_to_head(sd: Dict[str, torch.Tensor], *, num_channels: int, im_size: int, patch_size: int,
                   depth: int) -> Dict[str, torch.Tensor]:
    """ViT_UNet (README) state_dict -> HViT_UNet (HEAD) state_dict.  Exact iff the PatchEncoder conv is the identity
    (HEAD never applies it); the shared LN is duplicated into LN1 and LN2."""
    out = {}
    p_fine = patch_size // 2 ** depth
    for k, v in sd.items():
        if k.startswith("PE.conv2d."):
            continue
        if k == "PE.position_embedding.weight":
            out[k] = _repatch_table(v, num_channels, im_size, p_fine, patch_size)
        elif ".LN." in k:
            out[k.replace(".LN.", ".LN1.")] = v.clone()
            out[k.replace(".LN.", ".LN2.")] = v.clone()
        else:
            out[k] = v.clone()
    if "PE.conv2d.weight" in sd:
        w, b = sd["PE.conv2d.weight"], sd["PE.conv2d.bias"]
        ident = torch.zeros_like(w)
        for c in range(w.shape[0]):
            ident[c, c, 1, 1] = 1.0
        if not (torch.equal(w, ident) and torch.count_nonzero(b) == 0):
            raise ValueError("README checkpoint applies a non-identity PatchEncoder conv; HViT_UNet has no such layer "
                             "(model.py:84-91), so the conversion would change the function")
    return out


def head_to_readme(sd: Dict[str, torch.Tensor], *, num_channels: int, im_size: int, patch_size: int,
                   depth: int) -> Dict[str, torch.Tensor]:
    """HViT_UNet (HEAD) state_dict -> ViT_UNet (README) state_dict.  Exact iff LN1 == LN2 in every block; the
    PatchEncoder conv is created as the identity."""
    out = {}
    p_fine = patch_size // 2 ** depth
    for k, v in sd.items():
        if k == "PE.position_embedding.weight":
            out[k] = _repatch_table(v, num_channels, im_size, patch_size, p_fine)
        elif ".LN1." in k:
            other = sd[k.replace(".LN1.", ".LN2.")]
            if not torch.equal(v, other):
                raise ValueError(f"{k}: LN1 and LN2 differ; the README variant shares one LayerNorm per block")
            out[k.replace(".LN1.", ".LN.")] = v.clone()
        elif ".LN2." in k:
            continue
        else:
            out[k] = v.clone()
    C = num_channels
    w = torch.zeros(C, C, 3, 3, dtype=sd["conv2d.weight"].dtype, device=sd["conv2d.weight"].device)
    for c in range(C):
        w[c, c, 1, 1] = 1.0
    out["PE.conv2d.weight"] = w
    out["PE.conv2d.bias"] = torch.zeros(C, dtype=w.dtype, device=w.device)
    return out


# ------------------------------------------------------------------ best-checkpoint.bin (run_denoising.py:88,100)
# The reference trains through benatools' TorchFitterBase (not vendored, not installable here): `fitter.fit(...)`
# writes `<folder>/best-checkpoint.bin` and `fitter.load(path)` restores it (run_denoising.py:84-100).  The file is a
# torch.save'd dict; the keys below are the ones that class is known to write (model / optimizer / scheduler state,
# best loss, epoch).  load_checkpoint also accepts a bare state_dict, and either LayerNorm layout (README <-> HEAD).
def save_checkpoint(path, model, optimizer=None, scheduler=None, best_summary_loss=None, epoch=0, sync=True):
    """Write a best-checkpoint.bin.  `model` may be a vit_unet_b200.dp.DataParallel wrapper: with sync=True EVERY rank
    must call this (the BatchNorm running statistics of rank 0 are broadcast first), and only rank 0 writes."""
    import torch.distributed as dist
    from .dp import DataParallel
    rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
    if isinstance(model, DataParallel) and sync:
        model.sync_buffers(0)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if rank != 0:
        return None
    blob = {"model_state_dict": sd,
            "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
            "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None,
            "best_summary_loss": best_summary_loss, "epoch": epoch}
    torch.save(blob, path)
    return path


def load_checkpoint(path, model, optimizer=None, scheduler=None, geometry=None, map_location="cpu"):
    """Restore a best-checkpoint.bin (or a bare state_dict) into `model` (plain or DataParallel-wrapped).  When the file
    holds the OTHER variant's layout (shared LN vs LN1/LN2) it is converted with readme_to_head / head_to_readme; pass
    geometry=dict(num_channels=, im_size=, patch_size=, depth=) for that.  Returns the checkpoint dict."""
    blob = torch.load(path, map_location=map_location, weights_only=False)
    sd = blob["model_state_dict"] if isinstance(blob, dict) and "model_state_dict" in blob else blob
    want = set(model.state_dict().keys())
    if set(sd.keys()) != want:
        if geometry is None:
            raise ValueError("checkpoint layout differs from the model's (README vs HEAD variant?); pass geometry= to convert")
        sd = head_to_readme(sd, **geometry) if any(".LN1." in k for k in sd) else readme_to_head(sd, **geometry)
    model.load_state_dict(sd)
    if optimizer is not None and isinstance(blob, dict) and blob.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(blob["optimizer_state_dict"])
    if scheduler is not None and isinstance(blob, dict) and blob.get("scheduler_state_dict") is not None:
        scheduler.load_state_dict(blob["scheduler_state_dict"])
    return blob
