"""GPU input pipeline for the denoising / segmentation step.

The reference prepares every sample on the host inside ``DataLoader(num_workers=2)`` workers: ``cv2.imread`` ->
``cv2.resize(im_size)`` -> albumentations ``ShiftScaleRotate(shift 0.2, scale 0.2, rotate 20, BORDER_CONSTANT)`` (train
only) -> ``Normalize(mean 0.456, std 0.224, max_pixel_value 255)`` -> ``/255`` -> CHW float
(vit_unet/torch/dataset.py:56-70, run_denoising.py:52-58) -- which starves a B200 long before the model does.
Here the decoded uint8 HWC batch is copied to the device once and everything after ``imread`` runs there, per batch,
in two kernels (vu_input.cu).  Each data-parallel rank runs it on its own shard; no collective is involved.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import ops


class DenoisingBatchPipeline:
    """``x, y = pipe(noisy_u8, clean_u8)`` with uint8 (B,H,W,C) CUDA tensors -> float32 (B,C,im,im) model input / target.

    x = ((u8/255 - mean)/std)/255 and y = u8/255, exactly the quirk of the reference (Normalize, then the dataset
    divides by 255 again; the mask is not normalised).  train=True draws one ShiftScaleRotate per sample (the same
    affine map for image and mask: bilinear for the image, nearest for the mask, zero border)."""

    def __init__(self, im_size: int = 224, train: bool = True, mean: float = 0.456, std: float = 0.224,
                 shift_limit: float = 0.2, scale_limit: float = 0.2, rotate_limit: float = 20.0, seed: Optional[int] = None):
        self.im_size, self.train, self.mean, self.std = im_size, train, mean, std
        self.shift_limit, self.scale_limit, self.rotate_limit = shift_limit, scale_limit, rotate_limit
        self.gen = torch.Generator()
        if seed is not None:
            self.gen.manual_seed(seed)

    def sample_affine(self, B: int) -> torch.Tensor:
        """(B, 6) float32 maps from OUTPUT to SOURCE pixel coordinates (the inverse of the cv2.warpAffine matrix built by
        cv2.getRotationMatrix2D(centre, angle, scale) plus the shift, as albumentations' shift_scale_rotate does)."""
        S = self.im_size
        u = torch.rand(B, 4, generator=self.gen, dtype=torch.float64) * 2 - 1
        ang = u[:, 0] * self.rotate_limit * math.pi / 180.0
        sc = 1.0 + u[:, 1] * self.scale_limit
        dx, dy = u[:, 2] * self.shift_limit * S, u[:, 3] * self.shift_limit * S
        cx = cy = S / 2.0
        a, b = sc * torch.cos(ang), sc * torch.sin(ang)
        # forward (source -> output):  [a b (1-a) cx - b cy + dx ; -b a b cx + (1-a) cy + dy]
        fwd = torch.zeros(B, 3, 3, dtype=torch.float64)
        fwd[:, 0, 0], fwd[:, 0, 1], fwd[:, 0, 2] = a, b, (1 - a) * cx - b * cy + dx
        fwd[:, 1, 0], fwd[:, 1, 1], fwd[:, 1, 2] = -b, a, b * cx + (1 - a) * cy + dy
        fwd[:, 2, 2] = 1.0
        inv = torch.linalg.inv(fwd)[:, :2, :].reshape(B, 6)
        return inv.float()

    def __call__(self, noisy_u8: torch.Tensor, clean_u8: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        S = self.im_size
        B = noisy_u8.shape[0]
        if noisy_u8.shape[1] != S or noisy_u8.shape[2] != S:
            noisy_u8, clean_u8 = ops.resize_u8hwc(noisy_u8, S, S), ops.resize_u8hwc(clean_u8, S, S)
        mats = self.sample_affine(B).to(noisy_u8.device, non_blocking=True) if self.train else None
        x = ops.warp_u8hwc_to_chw(noisy_u8, mats, S, S, bilinear=True, scale=1.0 / 255.0, mean=self.mean, std=self.std,
                                  post=1.0 / 255.0)
        y = ops.warp_u8hwc_to_chw(clean_u8, mats, S, S, bilinear=False, scale=1.0 / 255.0)
        return x, y
