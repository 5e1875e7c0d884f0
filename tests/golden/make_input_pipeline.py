#!/usr/bin/env python
"""Generates tests/golden/input_pipeline.npz with OpenCV (needs cv2; run in the build container):
cv2.resize + cv2.getRotationMatrix2D + cv2.warpAffine on a small synthetic uint8 batch -- the calls made by
DenoisingDataset.__getitem__ (dataset.py:59-60) and albumentations' ShiftScaleRotate (run_denoising.py:52-55)."""
import os

import cv2
import numpy as np

rng = np.random.RandomState(7)
B, Hs, Ws, S = 3, 37, 53, 32
yy, xx = np.mgrid[0:Hs, 0:Ws]
imgs = np.stack([((np.sin(xx * (0.2 + 0.1 * b)) + np.cos(yy * (0.15 + 0.05 * b))) * 60 + 128 + rng.randint(-20, 20, (Hs, Ws)))[..., None]
                 * np.array([1.0, 0.8, 0.6]) for b in range(B)]).clip(0, 255).astype(np.uint8)
resized = np.stack([cv2.resize(im, (S, S)) for im in imgs])
params = [(12.0, 1.1, 3.0, -2.0), (-18.0, 0.85, -4.5, 5.0), (5.0, 1.0, 0.0, 0.0)]
fwd, warped_lin, warped_nn = [], [], []
for im, (ang, sc, dx, dy) in zip(resized, params):
    M = cv2.getRotationMatrix2D((S / 2, S / 2), ang, sc); M[0, 2] += dx; M[1, 2] += dy
    fwd.append(M)
    warped_lin.append(cv2.warpAffine(im, M, (S, S), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
    warped_nn.append(cv2.warpAffine(im, M, (S, S), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "input_pipeline.npz"), imgs=imgs, resized=resized,
                    fwd=np.stack(fwd), warped_lin=np.stack(warped_lin), warped_nn=np.stack(warped_nn))
