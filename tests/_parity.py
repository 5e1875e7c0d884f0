"""Shared parity protocol for the tensor-core (TF32 / bf16-map / streamed) paths against the golden vectors that
tests/golden/make_golden.py produced by executing the reference's own model.py.

Every tensor T (eval output, eval-mode and train-mode outputs / loss / dx / every parameter gradient, BN buffers) is
compared as  |T_cuda - T_ref|_max / |T_ref|_max <= base + YARD * cond(T)  where
  * T_ref is the reference evaluated in fp64 ("r64:" keys) where stored, else its fp32 evaluation;
  * cond(T) is the distance of the reference's OWN fp32 evaluation from that fp64 evaluation (the conditioning
    yardstick of test_model_gpu.py);
  * base is the precision class of the path under test: 1e-2 for outputs (north_star's TF32/BF16 bar), GRAD_BASE for
    gradients (products of two tensor-core-rounded factors summed over O(1e5) terms).
A tensor whose yardstick alone exceeds CHAOS is not reproducible by ANY fp32 implementation (Base train mode,
depth 2: dx 2e-1); it is reported, checked for finiteness and for a bounded relative L2 distance only.
"""
import os

import numpy as np
import torch

from make_golden import CONFIGS, fill_state_dict, make_input, pack, run_case

GOLD = os.path.join(os.path.dirname(__file__), "golden")
YARD = 10
CHAOS = 1e-2
OUT_BASE = 1e-2
GRAD_BASE = 5e-2          # TF32 contractions, fp32 attention maps: measured <= 2.8e-2 (B200, profiles/r02_parity.md); the margin
                          #   covers the run-to-run spread of the atomically accumulated reductions
GRAD_BASE_BF16 = 2.5e-1   # + bf16 storage of the mixed / gradient maps (2^-9 per element) on the TINY configs, whose BatchNorm
                          #   normalises over a few hundred map values (16 / 64 tokens, batch 2-3) and whose q/k conv gradients
                          #   are differences of nearly equal terms: measured 9.7e-2 .. 1.3e-1 (run-to-run: atomics order)
GRAD_BASE_L2 = 1e-2       # the Base / Lite / 1-channel LEVEL-2 shapes the benchmark spends its time at (l2block_* configs,
                          #   784 tokens): measured <= 3e-3 with either map storage, streamed or materialised
OUT_BASE_TRAIN_BF16 = 5e-2   # TRAIN-mode outputs with bf16 maps: BatchNorm batch statistics of the tiny configs are taken over as
                             #   few as 2 x 4 x 4 map values per head; measured <= 2.0e-2 (tiny_head_1ch), <= 1.7e-3 at Base shapes
CHAOS_TC = 1e-4           # a tensor whose fp32 reference is itself > 1e-4 from fp64 amplifies the 2^-11 input rounding of a
                          #   tensor-core path by the same factor (x 8192) to O(1): reported, finiteness only


class _Wrap(torch.nn.Module):        # run_case drives a CPU-style module; hop to the device at the boundary
    def __init__(self, m):
        super().__init__(); self.m = m

    def forward(self, t):
        return self.m(t.cuda()).cpu()


def build_net(name, quiet):
    import vit_unet_b200 as vu
    variant, kw, B = CONFIGS[name]
    net = quiet(vu.HViT_UNet, **kw)
    net.load_state_dict(fill_state_dict(net.state_dict()))
    net.to("cuda")
    x, y = make_input(B, kw["num_channels"], kw["im_size"])
    return net, x, y


def parity_rows(name, net, x, y, l1_grads=False, grad_base=GRAD_BASE, chaos=CHAOS / YARD, train_out_base=OUT_BASE):
    """[(key, err, tol, status)] with status in {'ok', 'FAIL', 'chaotic'}; err relative to max|ref|.

    Two passes over the golden file: the L1-loss run (tags evg / trn: the benchmark's loss) contributes the eval output,
    the eval- and train-mode outputs, losses and the BatchNorm buffers -- and its gradients only with l1_grads=True (the
    FP32 path), because d|e|/de = sign(e) is discontinuous: a path that is 1e-3 away in the OUTPUT flips the sign of
    the residual at a few pixels and per-pixel gradient tensors then differ by whole terms; the MSE-loss run (tags
    mev / mtr: the loss the reference trains with, run_denoising.py:80) contributes everything, gradients included."""
    gold = np.load(os.path.join(GOLD, f"{name}.npz"))
    cond = {k: float(gold[k]) for k in gold.files if "_cond:" in k}
    got = {}
    net.load_state_dict(fill_state_dict(net.state_dict()))      # fresh BN running statistics
    got.update(pack(run_case(_Wrap(net), x, y, train=True), full=name.startswith("tiny")))
    net.load_state_dict(fill_state_dict(net.state_dict()))
    got.update({k: v for k, v in pack(run_case(_Wrap(net), x, y, train=True, loss_kind="mse"), False).items()
                if k.startswith(("mev_", "mtr_"))})
    rows = []
    for k in gold.files:
        if k == "n_params" or "_cond:" in k or k.endswith("_sum") or "_gnorm:" in k or k.startswith("r64:"):
            continue
        l1_tag = k.startswith(("evg_", "trn_"))
        if l1_tag and not l1_grads and k[4:] not in ("out", "loss"):
            continue
        kk = k
        for tag in ("evg_g:", "trn_g:", "mev_g:", "mtr_g:", "buf:"):
            if k.startswith(tag):
                kk = tag + "m." + k[len(tag):]
        ref = gold["r64:" + k] if ("r64:" + k) in gold.files else gold[k]
        o = np.asarray(got[kk], dtype=np.float64)
        ref = np.asarray(ref, dtype=np.float64)
        if k.endswith("num_batches_tracked"):
            rows.append((k, float(np.abs(o - ref).max()), 0.0, "ok" if np.array_equal(o, ref) else "FAIL"))
            continue
        scale = max(float(np.abs(ref).max()), 1e-30)
        err = float(np.abs(o - ref).max()) / scale
        if k == "eval_out":
            c, base = 0.0, OUT_BASE
        elif k.startswith("buf:"):
            c, base = cond["trn_cond:out"], train_out_base
        else:
            tag, rest = k[:3], k[4:]
            if rest in ("out", "loss"):
                c, base = cond[f"{tag}_cond:out"], (train_out_base if tag in ("trn", "mtr") else OUT_BASE)
            elif rest == "dx":
                c, base = cond[f"{tag}_cond:dx"], grad_base
            else:
                pname = rest[2:]
                if tag in ("trn", "mtr") and pname.endswith("reatten_matrix.bias"):
                    continue          # exactly 0 in theory under train-mode BN; round-off on both sides
                c, base = cond[f"{tag}_cond:{pname}"], grad_base
        tol = base + YARD * c
        if not np.isfinite(o).all():
            rows.append((k, float("inf"), tol, "FAIL"))
        elif c > chaos:
            rows.append((k, err, tol, "chaotic"))
        else:
            rows.append((k, err, tol, "ok" if err <= tol else "FAIL"))
    return rows


def summarize(rows, top=8):
    worst = sorted((r for r in rows if r[3] != "chaotic"), key=lambda r: -r[1] / max(r[2], 1e-30))[:top]
    return "; ".join(f"{k}: {e:.2e}/{t:.1e}" for k, e, t, _ in worst)
