"""CPU oracle for the ViT-UNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file.  The shipped path (``vit_unet_b200``) never does; it fails loudly when its
CUDA extension is missing.

This is a plain-PyTorch fp32 *restatement* of the reference algorithm, written from the math of
``vit_unet/torch/model.py`` (reference cites below are relative to the upstream repo root):

* layout functions      model.py:8-53     (patch / unflatten / unpatch / downsampling / upsampling)
* PatchEncoder          model.py:57-91    (V-HEAD)   and  ViT_UNet.ipynb c16 (V-README / V-NB)
* FeedForward           model.py:95-110
* ReAttention           model.py:113-164
* encoder block         model.py:167-207  (V-HEAD: LN1/LN2)  and  ViT_UNet.ipynb c27 (shared LN)
* SkipConnection        model.py:211-259
* HViT_UNet             model.py:263-435, presets model.py:438-486
* ViT_UNet (README)     README.md:18-67, ViT_UNet.ipynb c44

Parity status: PINNED.  ``tests/golden/make_golden.py`` executes the *reference's own* model.py (read
from /root/reference at generation time, with the two constructor fixes SURVEY.md section 8(c) documents
applied in memory) on seeded weights/inputs and commits the outputs under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those vectors and against the reference's
published parameter counts (README.md:16,34,52) and shape known-answers (ViT_UNet.ipynb c47).

Differences from the reference that do not change results:
* patchify / unpatchify are single reshape+permute ops instead of unfold / Python cat loops
  (the reference's own functions are pure permutations: model.py:16-17, :33-34);
* q/k/v convs run on a (B*N, C, p, p) batch instead of a Python loop over B (model.py:152-154);
* the position ids are created on the input's device instead of being pinned to cuda:0 (model.py:71-75);
* ``torchvision.transforms.Resize`` (model.py:376) is an identity at matching sizes and is replaced by a
  shape assertion.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

__all__ = [
    "patchify", "unpatchify", "resample", "HViT_UNet", "ViT_UNet", "get_vit_unet",
    "dice_loss", "PRESETS",
]


# --------------------------------------------------------------------------------------- layout
def patchify(img: torch.Tensor, p: int) -> torch.Tensor:
    """(B,C,H,W) -> (B,N,C*p*p); token r*G+c, feature ch*p*p+i*p+j  (model.py:8-18 + flatten :86)."""
    B, C, H, W = img.shape
    assert H % p == 0, "Patch size must divide images height"
    assert W % p == 0, "Patch size must divide images width"
    gh, gw = H // p, W // p
    t = img.reshape(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5)
    return t.reshape(B, gh * gw, C * p * p)


def unpatchify(tok: torch.Tensor, C: int) -> torch.Tensor:
    """(B,N,C*p*p) -> (B,C,G*p,G*p) with a square token grid (model.py:20-35)."""
    B, N, D = tok.shape
    p = int(math.isqrt(D // C))
    g = int(math.isqrt(N))
    assert g * g == N and C * p * p == D
    t = tok.reshape(B, g, g, C, p, p).permute(0, 3, 1, 4, 2, 5)
    return t.reshape(B, C, g * p, g * p)


def resample(tok: torch.Tensor, C: int, factor: float) -> torch.Tensor:
    """downsampling (factor=0.5, model.py:39-45) / upsampling (factor=2, model.py:47-53)."""
    D = tok.shape[-1]
    p = int(math.isqrt(D // C))
    return patchify(unpatchify(tok, C), int(p * factor))


# --------------------------------------------------------------------------------------- blocks
def _patch_conv(x: torch.Tensor, conv: nn.Conv2d, C: int) -> torch.Tensor:
    """3x3 conv on every patch as a (C,p,p) image, zero padded at patch borders (model.py:152)."""
    B, N, D = x.shape
    p = int(math.isqrt(D // C))
    y = conv(x.reshape(B * N, C, p, p))
    return y.reshape(B, N, D)


class _ReAttnCore(nn.Module):
    """Parameters + math shared by ReAttention (model.py:113-164) and SkipConnection (:211-259)."""

    def __init__(self, dim, num_channels, num_heads, attn_drop, proj_drop, ksize=3):
        super().__init__()
        self.num_heads = num_heads
        self.num_channels = num_channels
        self.scale = (dim // num_heads) ** -0.5
        self.reatten_matrix = nn.Conv2d(num_heads, num_heads, 1, 1)
        self.var_norm = nn.BatchNorm2d(num_heads)
        self.qconv2d = nn.Conv2d(num_channels, num_channels, ksize, padding="same", bias=False)
        self.kconv2d = nn.Conv2d(num_channels, num_channels, ksize, padding="same", bias=False)
        self.vconv2d = nn.Conv2d(num_channels, num_channels, ksize, padding="same", bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def attend(self, xq, xk, xv):
        B, N, D = xq.shape
        h = self.num_heads
        C = self.num_channels

        def heads(t):
            return t.reshape(B, N, h, D // h).permute(0, 2, 1, 3)

        q = heads(_patch_conv(xq, self.qconv2d, C))
        k = heads(_patch_conv(xk, self.kconv2d, C))
        v = heads(_patch_conv(xv, self.vconv2d, C))
        attn = torch.matmul(q, k.transpose(-2, -1)) * self.scale
        attn = F.softmax(attn, dim=-1)
        attn = self.attn_drop(attn)
        attn = self.var_norm(self.reatten_matrix(attn))  # reatten_scale == 1.0 (model.py:140)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(B, N, D)
        return self.proj_drop(self.proj(out))


class ReAttention(_ReAttnCore):
    def forward(self, x):
        return self.attend(x, x, x)


class SkipConnection(_ReAttnCore):
    def forward(self, q, k, v):
        assert q.shape == k.shape
        assert k.shape == v.shape
        return self.attend(q, k, v)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden, dropout, dtype=None):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(dim, hidden, dtype=dtype), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden, dim, dtype=dtype), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class EncoderBlock(nn.Module):
    """Post-norm block with LayerNorm over the whole (N,D) token matrix.

    shared_ln=False -> V-HEAD  (LN1, LN2;  model.py:193-207)
    shared_ln=True  -> V-README (one LN used twice;  ViT_UNet.ipynb c27)
    """

    def __init__(self, N, C, D, hidden, heads, attn_drop, proj_drop, linear_drop, shared_ln, dtype=None):
        super().__init__()
        self.shared_ln = shared_ln
        self.ReAttn = ReAttention(D, C, heads, attn_drop, proj_drop)
        if shared_ln:
            self.LN = nn.LayerNorm((N, D), dtype=dtype)
        else:
            self.LN1 = nn.LayerNorm((N, D))
            self.LN2 = nn.LayerNorm((N, D))
        self.FeedForward = FeedForward(D, hidden, linear_drop, dtype=dtype)

    def forward(self, x):
        ln1, ln2 = (self.LN, self.LN) if self.shared_ln else (self.LN1, self.LN2)
        x = ln1(self.ReAttn(x) + x)
        x = ln2(self.FeedForward(x) + x)
        return x


class PatchEncoderHead(nn.Module):
    """V-HEAD: tokens + table[(N, D)]; the declared conv is never applied (model.py:79 vs :84-91)."""

    def __init__(self, im_size, p, C):
        super().__init__()
        self.p, self.C = p, C
        self.num_patches = (im_size // p) ** 2
        self.position_embedding = nn.Embedding(self.num_patches, C * p * p)

    def forward(self, X):
        return patchify(X, self.p) + self.position_embedding.weight


class PatchEncoderReadme(nn.Module):
    """V-README / V-NB: conv3x3(X), table indexed at the finest patch size (ViT_UNet.ipynb c16)."""

    def __init__(self, depth, num_patches, p, C, preprocessing):
        super().__init__()
        assert preprocessing in ["conv", "fourier", "none"], \
            "Preprocessing can only be 'conv', 'fourier' or 'none'."
        self.p, self.C, self.preprocessing = p, C, preprocessing
        self.p_final = p // (2 ** depth)
        self.n_final = num_patches * (4 ** depth)
        if preprocessing == "conv":
            self.conv2d = nn.Conv2d(C, C, 3, padding="same")
        self.position_embedding = nn.Embedding(self.n_final, C * self.p_final ** 2)

    def forward(self, X):
        if self.preprocessing == "conv":
            X = self.conv2d(X)
        elif self.preprocessing == "fourier":
            X = torch.fft.fft2(X).real
        fine = patchify(X, self.p_final) + self.position_embedding.weight
        return patchify(unpatchify(fine, self.C), self.p)


class _UNetBase(nn.Module):
    def _build(self, depth, depth_te, size_bottleneck, N0, D0, C, hidden, heads,
               attn_drop, proj_drop, linear_drop, shared_ln, dtype=None):
        def block(level):
            return EncoderBlock(N0 * 4 ** level, C, D0 // 4 ** level, hidden // 2 ** level, heads,
                                attn_drop, proj_drop, linear_drop, shared_ln, dtype)

        self.Encoders = nn.ModuleList(block(l) for l in range(depth) for _ in range(depth_te))
        self.BottleNeck = nn.ModuleList(block(depth) for _ in range(size_bottleneck))
        self.Decoders = nn.ModuleList()
        self.SkipConnections = nn.ModuleList()
        for level in range(depth):
            for _ in range(depth_te):
                self.Decoders.append(block(depth - level))
            self.SkipConnections.append(
                SkipConnection(D0 // 4 ** (depth - level - 1), C, heads, attn_drop, proj_drop))

    def _trunk(self, x):
        C, te, depth = self.num_channels, self.depth_te, self.depth
        skips = []
        for i, enc in enumerate(self.Encoders):
            x = enc(x)
            if (i + 1) % te == 0:
                skips.append(x)
                x = resample(x, C, 0.5)
        for b in self.BottleNeck:
            x = b(x)
        for i, dec in enumerate(self.Decoders):
            x = dec(x)
            if (i + 1) % te == 0:
                x = resample(x, C, 2)
                lvl = (i + 1) // te
                enc_x = skips[depth - lvl]
                assert enc_x.shape == x.shape, "enc and dec not same shape"
                x = self.SkipConnections[lvl - 1](enc_x, x, x)       # model.py:418
        return x


class HViT_UNet(_UNetBase):
    """V-HEAD (model.py:263-435) with the two constructor fixes of SURVEY.md section 8(c)."""

    def __init__(self, depth, depth_te, size_bottleneck, preprocessing, im_size, patch_size,
                 num_channels, hidden_dim, num_heads, attn_drop, proj_drop, linear_drop, verbose=False):
        super().__init__()
        assert patch_size % (2 ** depth) == 0, "Depth must be adjusted, final patch size is incompatible."
        assert patch_size // (2 ** depth) >= 4, \
            "Depth must be adjusted, final patch size is too small (lower than 4)."
        assert im_size % patch_size == 0, "Patch size is not compatible with image size."
        self.depth, self.depth_te, self.size_bottleneck = depth, depth_te, size_bottleneck
        self.preprocessing, self.im_size, self.patch_size = preprocessing, im_size, patch_size
        self.num_channels = num_channels
        self.num_patches = (im_size // patch_size) ** 2
        self.projection_dim = num_channels * patch_size ** 2
        self.PE = PatchEncoderHead(im_size, patch_size, num_channels)
        self._build(depth, depth_te, size_bottleneck, self.num_patches, self.projection_dim, num_channels,
                    hidden_dim, num_heads, attn_drop, proj_drop, linear_drop, shared_ln=False)
        if preprocessing == "conv":
            self.conv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same")

    def forward(self, X):
        assert X.shape[-1] == self.im_size and X.shape[-2] == self.im_size
        B = X.shape[0]
        x = self._trunk(self.PE(X))
        out = unpatchify(x, self.num_channels).reshape(B, self.num_channels, self.im_size, self.im_size)
        if self.preprocessing == "conv":
            out = self.conv2d(out)
        return out


class ViT_UNet(_UNetBase):
    """V-README (README.md:18-67): shared LN, PE conv applied, finest-granularity position table."""

    def __init__(self, depth, depth_te, size_bottleneck, preprocessing, num_patches, patch_size,
                 num_channels, hidden_dim, num_heads, attn_drop, proj_drop, linear_drop,
                 dtype=torch.float32):
        super().__init__()
        assert patch_size % (2 ** depth) == 0, "Depth must be adjusted, final patch size is incompatible."
        assert patch_size // (2 ** depth) >= 4, \
            "Depth must be adjusted, final patch size is too small (lower than 4)."
        self.depth, self.depth_te, self.size_bottleneck = depth, depth_te, size_bottleneck
        self.preprocessing, self.num_patches, self.patch_size = preprocessing, num_patches, patch_size
        self.num_channels = num_channels
        self.projection_dim = num_channels * patch_size ** 2
        self.im_size = int(math.isqrt(num_patches)) * patch_size
        self.PE = PatchEncoderReadme(depth, num_patches, patch_size, num_channels, preprocessing)
        self._build(depth, depth_te, size_bottleneck, num_patches, self.projection_dim, num_channels,
                    hidden_dim, num_heads, attn_drop, proj_drop, linear_drop, shared_ln=True, dtype=dtype)
        if preprocessing == "conv":
            self.conv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same")

    def forward(self, X):
        B, ch, H, W = X.shape
        x = self._trunk(self.PE(X))
        out = unpatchify(x, self.num_channels).reshape(B, ch, H, W)
        if self.preprocessing == "conv":
            out = self.conv2d(out)
        return out


PRESETS = {   # model.py:438-486 / README.md:18-67
    "lite": dict(depth=2, depth_te=1, size_bottleneck=2, patch_size=16, hidden_dim=64, num_heads=4),
    "base": dict(depth=2, depth_te=2, size_bottleneck=2, patch_size=32, hidden_dim=128, num_heads=8),
    "large": dict(depth=2, depth_te=4, size_bottleneck=4, patch_size=32, hidden_dim=128, num_heads=8),
}


def get_vit_unet(model_string: str, verbose=False, variant="head", num_channels=3,
                 attn_drop=0.2, proj_drop=0.2):
    key = model_string.lower()
    if key not in PRESETS:
        raise ValueError(f"Model string {model_string} not valid")
    cfg = PRESETS[key]
    common = dict(depth=cfg["depth"], depth_te=cfg["depth_te"], size_bottleneck=cfg["size_bottleneck"],
                  preprocessing="conv", patch_size=cfg["patch_size"], num_channels=num_channels,
                  hidden_dim=cfg["hidden_dim"], num_heads=cfg["num_heads"],
                  attn_drop=attn_drop, proj_drop=proj_drop, linear_drop=0)
    if variant == "head":
        return HViT_UNet(im_size=224, verbose=verbose, **common)
    return ViT_UNet(num_patches=(224 // cfg["patch_size"]) ** 2, **common)


def dice_loss(inp: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Soft Dice exactly as README.md:91-101 (smooth=1, whole batch flattened, raw outputs)."""
    smooth = 1.0
    i, t = inp.reshape(-1), target.reshape(-1)
    inter = (i * t).sum()
    return 1 - ((2.0 * inter + smooth) / (i.sum() + t.sum() + smooth))
