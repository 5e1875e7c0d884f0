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
