"""``vit_unet.torch.functions`` drop-in: the pieces of the reference module that touch the hot path.

``psnr`` keeps the reference contract (functions.py:7-19: no_grad, model(x) per batch, one PSNR per image, numpy
array out) but the per-image PSNR is computed by the ``vu_psnr`` CUDA kernel (same data-range rule as
``skimage.metrics.peak_signal_noise_ratio`` on float images) instead of copying every batch to the host.
"""
import numpy as np
import torch

from vit_unet_b200 import ops


def psnr(model, dataloader, data_range=None):
    score = []
    with torch.no_grad():
        for batch in dataloader:
            x = batch['x'].to('cuda').float()
            y = batch['y'].to('cuda').float().contiguous()
            out = model(x).contiguous()
            score.append(ops.psnr(out, y, 0.0 if data_range is None else float(data_range)).cpu())
    return np.asarray(torch.cat(score).numpy())
