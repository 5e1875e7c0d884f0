"""``vit_unet.torch.functions`` drop-in: the pieces of the reference module that touch the hot path.

``psnr`` keeps the reference contract (functions.py:7-19: no_grad, model(x) per batch, one PSNR per image,
numpy array out) but computes the per-image PSNR on the device with torch ops instead of skimage on the host.
"""
import numpy as np
import torch


def psnr(model, dataloader, data_range=None):
    score = []
    with torch.no_grad():
        for batch in dataloader:
            x = batch['x'].to('cuda').float()
            y = batch['y'].to('cuda').float()
            out = model(x)
            # skimage.metrics.peak_signal_noise_ratio: data_range from dtype (float -> 2.0 span [-1,1]) unless given
            dr = 2.0 if data_range is None else float(data_range)
            mse = ((out - y) ** 2).flatten(1).mean(dim=1)
            score.append((10.0 * torch.log10(dr * dr / mse)).cpu())
    return np.asarray(torch.cat(score).numpy())
