"""bf16 storage / compute mode (set_precision('bf16'), ViT_UNet(dtype=torch.bfloat16); BASELINE configs[3]):
the new kernel features it rests on -- MN-major bf16 B operands of the tcgen05 GEMM (weight gradients), the fused
epilogues with bf16 outputs / bf16 pre-activations, bf16 side outputs of LayerNorm, bf16 dropout / column sums, the
per-step weight casts -- against fp64 / fp32 PyTorch, and the whole model against the reference's golden vectors
(protocol: tests/_parity.py).  The residual stream, all statistics and all parameter gradients stay fp32."""
import contextlib
import io
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vit_unet_b200 import ops as _ops
    return _ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


def _close(a, b, tol, name=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 128), (128, 128, 256), (192, 192, 1568), (32, 192, 3000), (192, 32, 1568),
                                    (64, 768, 392), (3072, 128, 784), (48, 16, 640), (130, 200, 520)])
@pytest.mark.parametrize("ta", [True, False])
@pytest.mark.parametrize("split", [1, 3])
def test_tc_gemm_bf16_mn_major_b(ops, M, N, K, ta, split):
    """dW[M,N] (+)= dY^T X: bf16 A (MN- or K-major) x bf16 MN-major B (k rows of N contiguous values) -> fp32 C, with
    split-K atomics and accumulation -- the weight-gradient products of the bf16 mode."""
    def pad8(n): return (n + 7) // 8 * 8
    A = torch.zeros(K, pad8(M)) if ta else torch.zeros(M, pad8(K))
    if ta: A[:, :M] = _rand(K, M, seed=1)
    else: A[:, :K] = _rand(M, K, seed=1)
    Bm = torch.zeros(K, pad8(N)); Bm[:, :N] = _rand(K, N, seed=2)
    Ab, Bb = A.bfloat16(), Bm.bfloat16()
    Al = (Ab[:, :M].t() if ta else Ab[:, :K]).double()
    base = _rand(M, (N + 3) // 4 * 4, seed=3)
    exp = Al @ Bb[:, :N].double() + base[:, :N].double()
    out = base.clone().cuda()
    ops.gemm(Ab.cuda(), Bb.cuda(), out, M, N, K, trans_a=ta, trans_b=False, lda=A.shape[1], ldb=Bm.shape[1],
             ldc=out.shape[1], accumulate=True, split_k=split, precision=ops.PREC_TF32)
    _close(out[:, :N], exp, 2e-5 * math.sqrt(K) + 1e-5, name=f"bf16 MN-major B ta={ta}")
    assert torch.equal(out[:, N:].cpu(), base[:, N:])


@pytest.mark.parametrize("M,N,K", [(256, 128, 192), (300, 32, 192), (1568, 64, 768), (130, 192, 32), (784, 3072, 128)])
def test_tc_gemm_bf16_epilogues(ops, M, N, K):
    """Linear -> GELU (+ bf16 pre-activation) -> dropout with bf16 output; GELU' x (dY W) with a bf16 pre-activation input;
    bias + fp32 residual into an fp32 output: every epilogue the bf16 mode uses, from bf16 operands."""
    A, W = _rand(M, K, seed=1).bfloat16(), (_rand(N, K, seed=2) / math.sqrt(K)).bfloat16()
    bias, res = _rand(N, seed=3), _rand(M, N, seed=4)
    lin = A.double() @ W.double().t() + bias.double()
    # 1. GELU with aux_out, bf16 C
    act = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    pre = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), act, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, bias=bias.cuda(), act=ops.ACT_GELU,
             aux_out=pre, ldaux=N, precision=ops.PREC_TF32)
    _close(pre.float(), lin, 6e-3, "bf16 pre-activation")
    _close(act.float(), F.gelu(lin), 6e-3, "bf16 GELU output")
    # 2. the same with dropout: kept elements scaled, mask identical to the standalone kernel's
    p = 0.25
    actd = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), actd, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, bias=bias.cuda(), act=ops.ACT_GELU,
             drop_p=p, drop_seed=5, drop_stream=9, precision=ops.PREC_TF32)
    ones = torch.ones(M, N, device="cuda")
    mask = ops.dropout(ones, torch.empty_like(ones), p, 5, 9)
    _close(actd.float(), F.gelu(lin) * mask.cpu().double(), 8e-3, "bf16 GELU + dropout")
    # 3. GELU backward with bf16 aux_in, bf16 C
    dpre = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), dpre, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, act=ops.ACT_GELU_BWD, aux_in=pre, ldaux=N,
             precision=ops.PREC_TF32)
    t = pre.float().cpu().double().requires_grad_(True)
    F.gelu(t).sum().backward()
    _close(dpre.float(), (A.double() @ W.double().t()) * t.grad, 6e-3, "bf16 GELU backward")
    # 4. bias + dropout + fp32 residual -> fp32 C (the proj / second FeedForward product)
    y = torch.zeros(M, N, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), y, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, bias=bias.cuda(), residual=res.cuda(),
             drop_p=p, drop_seed=5, drop_stream=9, precision=ops.PREC_TF32)
    _close(y, lin * mask.cpu().double() + res.double(), 1e-4, "fp32 C from bf16 operands")


# ------------------------------------------------------------------------------------------------ elementwise
def test_layernorm_bf16_side_outputs(ops):
    B, N, D = 3, 49, 192
    n = N * D
    x, g = _rand(B, N, D, seed=1).cuda(), _rand(B, N, D, seed=2).cuda()
    w, b = (_rand(N, D, seed=3) + 1.5).cuda(), _rand(N, D, seed=4).cuda()
    st = torch.empty(B, 2, device="cuda")
    ops.ln_stats(x, B, n, 1e-5, st)
    out, out16 = torch.empty_like(x), torch.empty(B, N, D, dtype=torch.bfloat16, device="cuda")
    ops.ln_apply(x, st, w, b, out, B, n, out16=out16)
    ref = F.layer_norm(x, (N, D), w, b, 1e-5)
    _close(out, ref, 2e-5, "ln fp32")
    assert torch.equal(out16, out.bfloat16())
    dx, dx16 = torch.empty_like(x), torch.empty(B, N, D, dtype=torch.bfloat16, device="cuda")
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    scratch = torch.empty(B, ops.LN_SCRATCH, device="cuda")
    ops.ln_bwd(g, x, st, w, dx, dw, db, scratch, B, n, dx16=dx16)
    xr = x.clone().requires_grad_(True)
    F.layer_norm(xr, (N, D), w, b, 1e-5).backward(g)
    _close(dx, xr.grad, 5e-5, "ln bwd fp32")
    assert torch.equal(dx16, dx.bfloat16())


def test_dropout_colsum_cast_bf16(ops):
    n = 4 * 1000 + 3
    x = _rand(n, seed=1).cuda()
    a, b = torch.empty_like(x), torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ops.dropout(x, a, 0.3, 7, 2)
    ops.dropout(x, b, 0.3, 7, 2)
    assert torch.equal(b, a.bfloat16())                     # same mask, rounded once
    ops.dropout(x, b, 0.0, 0, 0)
    assert torch.equal(b, x.bfloat16())                     # p = 0: plain conversion
    M, N = 1000, 72
    X = _rand(M, N, seed=2).bfloat16().cuda()
    out = torch.zeros(N, device="cuda")
    ops.colsum(X, M, N, N, out)
    _close(out, X.double().sum(0), 1e-5, "colsum bf16")
    W = _rand(70, 130, seed=3).cuda()
    wn, wt = ops.cast_bf16(W)
    assert torch.equal(wn, W.bfloat16()) and torch.equal(wt, W.t().contiguous().bfloat16())
    wn2, wt2 = ops.cast_bf16(W, want_t=False)
    assert wt2 is None and torch.equal(wn2, wn)


# ------------------------------------------------------------------------------------------------ model
@pytest.fixture
def bf16():
    import vit_unet_b200 as vu
    vu.set_precision("bf16")
    yield vu
    vu.set_precision("fp32"); vu.set_bf16_maps(True); vu.set_streamed(False)


# bf16 operands round to 2^-9 (TF32: 2^-11): four times the operand rounding of the TF32 class.  Outputs keep north_star's
# 1e-2; gradients get the bf16-map bases of tests/_parity.py scaled accordingly where measured (profiles/r02_parity_bf16.md)
GRAD_BASE_L2_B16 = 2e-2
GRAD_BASE_B16 = 2.5e-1


@pytest.mark.parametrize("name", ["l2block_head", "l2block_1ch", "l2block_lite", "tiny_head", "lite_head", "base_head"])
def test_bf16_mode_matches_reference_golden(bf16, name):
    from _parity import CHAOS_TC, OUT_BASE_TRAIN_BF16, build_net, parity_rows, summarize
    net, x, y = build_net(name, _quiet)
    gb = GRAD_BASE_L2_B16 if name.startswith("l2block") else GRAD_BASE_B16
    rows = parity_rows(name, net, x, y, grad_base=gb, chaos=CHAOS_TC, train_out_base=OUT_BASE_TRAIN_BF16)
    bad = [r for r in rows if r[3] == "FAIL"]
    assert not bad, f"{len(bad)} of {len(rows)} tensors out of tolerance; worst: {summarize(bad)}"


def test_bf16_mode_eval_output_and_psnr():
    """north_star's acceptance for the BF16 tensor-core path: eval output within 1e-2 of the reference (oracle, same
    weights and inputs), PSNR delta < 0.01 dB -- Base and Large-shaped (depth_te = 4) README models built with
    dtype=torch.bfloat16, parameters staying fp32 masters."""
    import vit_unet_b200 as vu
    from make_golden import fill_state_dict, make_input
    from oracle import vit_unet_oracle as O
    for depth_te, bott in ((2, 2), (4, 4)):
        kw = dict(depth=2, depth_te=depth_te, size_bottleneck=bott, preprocessing="conv", num_patches=49, patch_size=32,
                  num_channels=3, hidden_dim=128, num_heads=8, attn_drop=0.2, proj_drop=0.2, linear_drop=0)
        ref = _quiet(O.ViT_UNet, **kw)
        net = _quiet(vu.ViT_UNet, dtype=torch.bfloat16, **kw)
        assert all(p.dtype == torch.float32 for p in net.parameters())
        sd = fill_state_dict(ref.state_dict())
        ref.load_state_dict(sd); net.load_state_dict(sd); net.to("cuda")
        ref.eval(); net.eval()
        x, clean = make_input(2, 3, 224)
        with torch.no_grad():
            a, b = ref(x), net(x.cuda()).cpu()
        rel = ((a - b).abs().max() / a.abs().max()).item()
        assert rel <= 1e-2, (depth_te, rel)
        psnr = lambda o: 10 * torch.log10(4.0 / ((o - clean) ** 2).flatten(1).mean(1))
        assert (psnr(a) - psnr(b)).abs().max().item() < 0.01, depth_te
        # a training step runs and yields finite fp32 gradients for every parameter
        net.train(); net.zero_grad()
        vu.l1_loss(net(x.cuda()), clean.cuda()).backward()
        assert all(p.grad is not None and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all() for p in net.parameters())


def test_bf16_mode_dropout_masks_agree_between_forward_and_backward(bf16):
    """Dropout on (attention 0.25 / projection 0.25 / FeedForward 0.25) in the bf16 mode: gradients against central
    differences of the CUDA forward under the same seed (same masks) at the level-2 block shape."""
    vu = bf16
    from _parity import CONFIGS
    from make_golden import fill_state_dict, make_input
    _, kw, _ = CONFIGS["l2block_head"]
    kw = dict(kw, attn_drop=0.25, proj_drop=0.25, linear_drop=0.25, size_bottleneck=1)
    net = _quiet(vu.HViT_UNet, **kw)
    net.load_state_dict(fill_state_dict(net.state_dict()))
    net.to("cuda").train()
    x, y = make_input(2, 3, 224)
    x, y = x.cuda(), y.cuda()

    def loss():
        torch.manual_seed(11)
        return vu.mse_loss(net(x), y)
    net.zero_grad(); loss().backward()
    pd = dict(net.named_parameters())
    for pname, idx in (("BottleNeck.0.ReAttn.proj.bias", 17), ("BottleNeck.0.FeedForward.net.3.bias", 40),
                       ("BottleNeck.0.FeedForward.net.0.bias", 5), ("BottleNeck.0.ReAttn.var_norm.weight", 3)):
        prm = pd[pname]
        g = prm.grad.view(-1)[idx].item()
        with torch.no_grad():
            eps = 5e-2 * max(1.0, abs(prm.view(-1)[idx].item()))
            prm.view(-1)[idx] += eps; lp = loss().item()
            prm.view(-1)[idx] -= 2 * eps; lm = loss().item()
            prm.view(-1)[idx] += eps
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - g) <= 0.15 * max(abs(g), abs(fd)) + 1e-4, (pname, fd, g)
