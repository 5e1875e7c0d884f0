"""Drop-in ``nn.Module`` surface of the reference's torch model, executed by the CUDA engine.

Mirrors (names, constructor signatures, state_dict keys/shapes, assertion messages):
  * ``HViT_UNet`` + ``get_vit_unet``  -- vit_unet/torch/model.py:263-486  (what run_denoising.py:78 calls)
  * ``ViT_UNet``                      -- README.md:18-67 / ViT_UNet.ipynb c44 (shared LN, PE conv, fine table)

The sub-modules below only HOLD parameters (so ``state_dict()`` keys, default initialisation and ``.to()`` /
optimizers behave exactly as with the reference); none of their ``forward`` methods is ever called.  The whole
network is one ``torch.autograd.Function`` whose forward/backward are kernel schedules (engine.py).
There is no CPU path: calling the model with a CPU tensor raises.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
from torch import nn

from . import ops
from .engine import Engine, Geometry


# ------------------------------------------------------------------------- parameter containers
class _AttnParams(nn.Module):
    """Parameter set of ReAttention (model.py:134-148) / SkipConnection (model.py:231-241)."""

    def __init__(self, dim, num_channels, num_heads, dtype=None):
        super().__init__()
        self.reatten_matrix = nn.Conv2d(num_heads, num_heads, 1, 1)
        self.var_norm = nn.BatchNorm2d(num_heads)
        self.qconv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same", bias=False)
        self.kconv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same", bias=False)
        self.vconv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same", bias=False)
        self.proj = nn.Linear(dim, dim)


class _FFParams(nn.Module):
    def __init__(self, dim, hidden, dropout, dtype=None):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden, dtype=dtype), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden, dim, dtype=dtype), nn.Dropout(dropout))


class _BlockParams(nn.Module):
    def __init__(self, N, C, D, hidden, heads, linear_drop, shared_ln, dtype=None):
        super().__init__()
        self.ReAttn = _AttnParams(D, C, heads)
        if shared_ln:
            self.LN = nn.LayerNorm((N, D), dtype=dtype)
        else:
            self.LN1 = nn.LayerNorm((N, D))
            self.LN2 = nn.LayerNorm((N, D))
        self.FeedForward = _FFParams(D, hidden, linear_drop, dtype)


class _PEParams(nn.Module):
    def __init__(self, n_table, d_table, C, conv):
        super().__init__()
        if conv:
            self.conv2d = nn.Conv2d(C, C, 3, padding="same")
        self.position_embedding = nn.Embedding(n_table, d_table)


# ------------------------------------------------------------------------- the single autograd node
class _ViTUNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, owner, names, *params):
        P = dict(zip(names, params))
        P.update(owner._buffer_dict())
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if owner.training else 0
        out, saved = owner.engine.forward(P, X, train=owner.training, save=True, seed=seed)
        ctx.owner, ctx.names, ctx.saved = owner, names, saved
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, dout):
        owner, names = ctx.owner, ctx.names
        params = ctx.saved_tensors
        P = dict(zip(names, params))
        P.update(owner._buffer_dict())
        flat = ops.zeros(owner._flat_numel, torch.float32, dout.device)      # cudaMemsetAsync, no fill kernel
        G = {n: flat[o:o + p.numel()].view(p.shape) for n, p, o in zip(names, params, owner._flat_offsets)}
        owner._last_flat_grad = flat
        dp = owner._dp
        if dp is not None:           # data parallel: all-reduce finished suffixes of `flat` while backward continues
            dp.begin(flat)
            owner.engine.on_grads_ready = dp.on_ready
        dX = owner.engine.backward(P, G, ctx.saved, dout.contiguous(), ctx.needs_input_grad[0])
        if dp is not None:
            dp.finish()
        ctx.saved = None
        return (dX, None, None) + tuple(G[n] for n in names)


class _ViTUNetBase(nn.Module):
    def _finish_init(self, geom: Geometry):
        self.engine = Engine(geom)
        names = [n for n, _ in self.named_parameters()]
        self._param_names: List[str] = geom.param_order(names)
        pd = dict(self.named_parameters())
        self._flat_offsets, off = [], 0
        for n in self._param_names:
            self._flat_offsets.append(off)
            off += (pd[n].numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self._flat_numel = off
        self._last_flat_grad = None
        self._dp = None

    def _build_blocks(self, depth, depth_te, size_bottleneck, N0, D0, C, hidden, heads, linear_drop, shared_ln,
                      dtype=None):
        def block(level):
            return _BlockParams(N0 * 4 ** level, C, D0 // 4 ** level, hidden // 2 ** level, heads, linear_drop,
                                shared_ln, dtype)
        self.Encoders = nn.ModuleList(block(l) for l in range(depth) for _ in range(depth_te))
        self.BottleNeck = nn.ModuleList(block(depth) for _ in range(size_bottleneck))
        self.Decoders = nn.ModuleList()
        self.SkipConnections = nn.ModuleList()
        for level in range(depth):
            for _ in range(depth_te):
                self.Decoders.append(block(depth - level))
            self.SkipConnections.append(_AttnParams(D0 // 4 ** (depth - level - 1), C, heads))

    def _buffer_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_buffers())

    def _check_input(self, X):
        g = self.engine.g
        if not X.is_cuda:
            raise RuntimeError("vit_unet_b200 runs on CUDA (sm_100a) only; move the model and the input to 'cuda'. "
                               "There is no CPU fallback.")
        assert X.dim() == 4 and X.shape[1] == g.C, "Num. channels must agree"
        assert X.shape[2] % g.p0 == 0, "Patch size must divide images height"
        assert X.shape[3] % g.p0 == 0, "Patch size must divide images width"

    def _run(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous().float()
        pd = dict(self.named_parameters())
        params = [pd[n] for n in self._param_names]
        needs_grad = torch.is_grad_enabled() and (X.requires_grad or any(p.requires_grad for p in params))
        if needs_grad:
            return _ViTUNetFn.apply(X, self, self._param_names, *params)
        P = {n: p.detach() for n, p in pd.items()}
        P.update(self._buffer_dict())
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) if self.training else 0
        B = X.shape[0]
        chunk = self._eval_chunk(B)
        if self.training or chunk >= B:
            out, _ = self.engine.forward(P, X, train=self.training, save=False, seed=seed)
            return out
        # inference: images are independent (BatchNorm uses running statistics), so the batch is processed in
        # slices that keep the transient attention maps (2 x B*h*N^2 floats at the finest level) within budget
        out = torch.empty_like(X)
        for s0 in range(0, B, chunk):
            o, _ = self.engine.forward(P, X[s0:s0 + chunk], train=False, save=False, seed=seed)
            out[s0:s0 + chunk] = o
        return out

    map_budget_bytes = 16 << 30       # transient attention-map budget for no-grad inference (HBM is 180 GB)

    def _eval_chunk(self, B: int) -> int:
        from . import engine as _e
        g = self.engine.g
        per_image = 1
        for l in range(g.depth + 1):         # levels whose maps are materialised (the streamed inference kernel writes none)
            n, hd = g.N(l), g.D(l) // g.heads
            if (self.engine._prec() == ops.PREC_TF32 and _e._STREAMED_INFER["value"] and ops.reattn_stream_supported(g.heads, hd, n)):
                continue
            per_image = max(per_image, 2 * g.heads * n * ((n + 3) // 4 * 4) * 4)
        return max(1, min(B, self.map_budget_bytes // per_image))

    def flat_grad(self):
        """The flat fp32 gradient buffer of the last backward (forward-execution parameter order)."""
        return self._last_flat_grad


def _print_arch(depth, patch_size, num_patches, projection_dim, hidden_dim):
    # the reference prints this table unconditionally at construction (model.py:301-307)
    print('Architecture information:')
    for i in range(depth + 1):
        print('Level {}:'.format(i))
        print('\tPatch size:', patch_size // (2 ** i))
        print('\tNum. patches:', num_patches * (4 ** i))
        print('\tProjection size:', projection_dim // (4 ** i))
        print('\tHidden dim. size:', hidden_dim // (2 ** i))


class HViT_UNet(_ViTUNetBase):
    """vit_unet/torch/model.py:263-435 (constructor defects at :78 and :309 fixed)."""

    def __init__(self, depth: int, depth_te: int, size_bottleneck: int, preprocessing: str, im_size: int,
                 patch_size: int, num_channels: int, hidden_dim: int, num_heads: int, attn_drop: float,
                 proj_drop: float, linear_drop: float, verbose: bool = False):
        super().__init__()
        assert patch_size % (2 ** depth) == 0, "Depth must be adjusted, final patch size is incompatible."
        assert patch_size // (2 ** depth) >= 4, "Depth must be adjusted, final patch size is too small (lower than 4)."
        assert im_size % patch_size == 0, "Patch size is not compatible with image size."
        if preprocessing == "fourier":
            raise NotImplementedError("preprocessing='fourier' discards the network output in the reference "
                                      "(model.py:429-430) and is out of scope")
        self.depth, self.depth_te, self.size_bottleneck = depth, depth_te, size_bottleneck
        self.preprocessing, self.im_size, self.patch_size = preprocessing, im_size, patch_size
        self.num_patches = (im_size // patch_size) ** 2
        self.num_channels = num_channels
        self.projection_dim = num_channels * patch_size ** 2
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.attn_drop, self.proj_drop, self.linear_drop = attn_drop, proj_drop, linear_drop
        self.verbose = verbose
        _print_arch(depth, patch_size, self.num_patches, self.projection_dim, hidden_dim)
        self.PE = _PEParams(self.num_patches, self.projection_dim, num_channels, conv=False)
        self._build_blocks(depth, depth_te, size_bottleneck, self.num_patches, self.projection_dim, num_channels,
                           hidden_dim, num_heads, linear_drop, shared_ln=False)
        if preprocessing == "conv":
            self.conv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same")
        self._finish_init(Geometry(C=num_channels, S=im_size, p0=patch_size, depth=depth, depth_te=depth_te,
                                   n_bottleneck=size_bottleneck, heads=num_heads, hidden=hidden_dim,
                                   shared_ln=False, pe_conv=False, table_p=patch_size,
                                   out_conv=(preprocessing == "conv"), attn_drop=attn_drop, proj_drop=proj_drop,
                                   linear_drop=linear_drop))

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        self._check_input(X)
        if X.shape[-1] != self.im_size or X.shape[-2] != self.im_size:
            # the reference resizes with torchvision (model.py:376); on the CUDA path inputs must already match
            raise AssertionError(f"input must be {self.im_size}x{self.im_size}; resize before calling the model")
        out = self._run(X)
        if self.verbose:
            print('Final')
            print(torch.cuda.memory_summary(X.device))
        return out


class ViT_UNet(_ViTUNetBase):
    """README.md:18-67 constructor; semantics of ViT_UNet.ipynb c16/c27/c44 with 3x3 q/k/v convs."""

    def __init__(self, depth: int, depth_te: int, size_bottleneck: int, preprocessing: str, num_patches: int,
                 patch_size: int, num_channels: int, hidden_dim: int, num_heads: int, attn_drop: float,
                 proj_drop: float, linear_drop: float, dtype: torch.dtype = torch.float32):
        super().__init__()
        assert patch_size % (2 ** depth) == 0, "Depth must be adjusted, final patch size is incompatible."
        assert patch_size // (2 ** depth) >= 4, "Depth must be adjusted, final patch size is too small (lower than 4)."
        assert preprocessing in ['conv', 'fourier', 'none'], "Preprocessing can only be 'conv', 'fourier' or 'none'."
        if preprocessing == "fourier":
            raise NotImplementedError("preprocessing='fourier' is out of scope (the reference discards the output)")
        if dtype not in (torch.float32, torch.bfloat16):
            raise NotImplementedError("dtype must be torch.float32 or torch.bfloat16")
        side = int(math.isqrt(num_patches))
        assert side * side == num_patches, "num_patches must be a perfect square"
        self.depth, self.depth_te, self.size_bottleneck = depth, depth_te, size_bottleneck
        self.preprocessing, self.num_patches, self.patch_size = preprocessing, num_patches, patch_size
        self.num_channels = num_channels
        self.projection_dim = num_channels * patch_size ** 2
        self.hidden_dim, self.num_heads = hidden_dim, num_heads
        self.attn_drop, self.proj_drop, self.linear_drop = attn_drop, proj_drop, linear_drop
        self.dtype = dtype
        self.im_size = side * patch_size
        _print_arch(depth, patch_size, num_patches, self.projection_dim, hidden_dim)
        p_final = patch_size // (2 ** depth)
        self.PE = _PEParams(num_patches * 4 ** depth, num_channels * p_final ** 2, num_channels,
                            conv=(preprocessing == "conv"))
        # dtype=torch.bfloat16 (the notebook forwards it to Linear / LayerNorm, ViT_UNet.ipynb c22:L10,13, c27:L27-29)
        # selects the bf16 storage / compute mode of the engine for THIS module: bf16 tcgen05 products and bf16 saved
        # activations, fp32 master parameters (so optimizers and checkpoints are unchanged), fp32 inputs and outputs
        self._build_blocks(depth, depth_te, size_bottleneck, num_patches, self.projection_dim, num_channels,
                           hidden_dim, num_heads, linear_drop, shared_ln=True, dtype=None)
        if preprocessing == "conv":
            self.conv2d = nn.Conv2d(num_channels, num_channels, 3, padding="same")
        self._finish_init(Geometry(C=num_channels, S=self.im_size, p0=patch_size, depth=depth, depth_te=depth_te,
                                   n_bottleneck=size_bottleneck, heads=num_heads, hidden=hidden_dim,
                                   shared_ln=True, pe_conv=(preprocessing == "conv"), table_p=p_final,
                                   out_conv=(preprocessing == "conv"), attn_drop=attn_drop, proj_drop=proj_drop,
                                   linear_drop=linear_drop))
        if dtype == torch.bfloat16:
            from .engine import PREC_BF16
            self.engine.precision = PREC_BF16

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        self._check_input(X)
        assert X.shape[-1] == self.im_size and X.shape[-2] == self.im_size, \
            f"input must be {self.im_size}x{self.im_size} (sqrt(num_patches)*patch_size)"
        return self._run(X)


_PRESETS = {   # model.py:438-486
    "lite": dict(depth=2, depth_te=1, size_bottleneck=2, patch_size=16, hidden_dim=64, num_heads=4),
    "base": dict(depth=2, depth_te=2, size_bottleneck=2, patch_size=32, hidden_dim=128, num_heads=8),
    "large": dict(depth=2, depth_te=4, size_bottleneck=4, patch_size=32, hidden_dim=128, num_heads=8),
}


def get_vit_unet(model_string: str, verbose=False):
    key = model_string.lower()
    if key not in _PRESETS:
        raise ValueError(f'Model string {model_string} not valid')
    c = _PRESETS[key]
    return HViT_UNet(depth=c["depth"], depth_te=c["depth_te"], size_bottleneck=c["size_bottleneck"],
                     preprocessing='conv', im_size=224, patch_size=c["patch_size"], num_channels=3,
                     hidden_dim=c["hidden_dim"], num_heads=c["num_heads"], attn_drop=.2, proj_drop=.2,
                     linear_drop=0, verbose=verbose)
