"""Per-kernel parity: every C-ABI entry point against the plain fp32 PyTorch op it replaces (run on CPU so
the check does not depend on cuBLAS/cuDNN TF32 settings).  Tolerances: 1e-5-class relative for the FP32 path."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import vit_unet_oracle as O          # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def ops():
    from vit_unet_b200 import ops as _ops
    return _ops


def _close(a, b, rtol=2e-5, atol_rel=2e-6, name=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(b.abs().max().item(), 1e-30)
    err = (a - b).abs().max().item()
    assert err <= rtol * scale + atol_rel * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g) * 2 - 1) * scale


# ------------------------------------------------------------------------------------------------ layout
@pytest.mark.parametrize("C,S,p_in,p_out", [(3, 32, 0, 16), (3, 32, 16, 8), (3, 32, 8, 16), (3, 64, 32, 0),
                                             (1, 64, 16, 4), (3, 28, 0, 7), (3, 28, 14, 7), (2, 24, 12, 0)])
def test_repatch(ops, C, S, p_in, p_out):
    B = 3
    img = _rand(B, C, S, S)
    src = img if p_in == 0 else O.patchify(img, p_in)
    exp = img if p_out == 0 else O.patchify(img, p_out)
    out = torch.empty(exp.shape, device="cuda")
    ops.repatch(src.contiguous().cuda(), out, B, C, S, S, p_in, p_out)
    assert torch.equal(out.cpu(), exp)            # pure permutation: bit exact


def test_pe_fwd_and_table_grad(ops):
    B, C, S, p, pt = 3, 3, 32, 16, 4
    img = _rand(B, C, S, S)
    table = _rand((S // pt) ** 2, C * pt * pt, seed=1)
    exp = O.patchify(O.unpatchify(O.patchify(img, pt) + table, C), p)
    out = torch.empty(exp.shape, device="cuda")
    ops.pe_fwd(img.cuda(), 0, table.cuda(), pt, out, p, B, C, S, S)
    assert torch.equal(out.cpu(), exp)
    dout = _rand(*exp.shape, seed=2)
    dt = torch.empty_like(table, device="cuda")
    ops.pe_bwd_table(dout.cuda(), p, dt, pt, B, C, S, S)
    exp_dt = O.patchify(O.unpatchify(dout, C), pt).sum(0)
    _close(dt, exp_dt, name="dtable")


# ------------------------------------------------------------------------------------------------ convs
@pytest.mark.parametrize("C,S,p,B", [(3, 32, 16, 2), (3, 32, 4, 2), (1, 32, 8, 2), (3, 28, 7, 2),
                                       (3, 64, 32, 3), (3, 24, 8, 3), (1, 20, 4, 2), (3, 48, 16, 3), (2, 32, 8, 1)])
@pytest.mark.parametrize("nconv", [1, 2, 3])
def test_patch_conv_fwd_bwd(ops, C, S, p, B, nconv):
    img = _rand(B, C, S, S)
    x = O.patchify(img, p).contiguous()
    N, D = x.shape[1], x.shape[2]
    ws = [_rand(C, C, 3, 3, seed=10 + k, scale=0.5) for k in range(nconv)]
    xr = x.clone().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    ys = [F.conv2d(xr.reshape(B * N, C, p, p), w, padding=1).reshape(B, N, D) for w in wr]
    outs = [torch.empty(B, N, D, device="cuda") for _ in range(nconv)]
    wcat = torch.cat([w.reshape(-1) for w in ws]).cuda()
    ops.conv3x3_fwd(x.cuda(), p, wcat, None, outs, p, p, B, C, S, S)
    for o, y in zip(outs, ys):
        _close(o, y, name="conv fwd")
    dys = [_rand(B, N, D, seed=20 + k) for k in range(nconv)]
    sum((y * d).sum() for y, d in zip(ys, dys)).backward()
    dx = torch.empty(B, N, D, device="cuda")
    ops.conv3x3_bwd_data([d.cuda() for d in dys], p, wcat, dx, p, p, B, C, S, S)
    _close(dx, xr.grad, name="conv dx")
    dw = torch.zeros(nconv * C * C * 9, device="cuda")
    db = torch.zeros(nconv * C, device="cuda")
    ops.conv3x3_bwd_weight(x.cuda(), p, [d.cuda() for d in dys], p, dw, db, p, B, C, S, S)
    _close(dw, torch.cat([w.grad.reshape(-1) for w in wr]), rtol=1e-4, name="conv dw")
    _close(db, torch.cat([d.reshape(B * N, C, p * p).sum((0, 2)) for d in dys]), rtol=1e-4, name="conv db")


@pytest.mark.parametrize("B,N,D,h", [(2, 784, 192, 8), (3, 196, 768, 8), (2, 49, 3072, 8), (2, 16, 64, 4), (2, 196, 48, 4),
                                     (1, 100, 36, 3), (2, 3136, 48, 4), (1, 70, 200, 5), (2, 9, 30, 3)])
def test_heads_transpose_bf16_bit_exact(ops, B, N, D, h):
    """(B,N,D) fp32 -> per-head transposed bf16 copy (K-major B operand of the map-reading GEMMs; head split of model.py:152):
    bit-equal to torch's round-to-nearest cast of the permuted tensor; the last shape (D % 4 != 0) takes the per-head kernel."""
    x = _rand(B, N, D, seed=5).cuda()
    out = ops.heads_transpose_bf16(x, B, N, D, h)
    ref = x.view(B, N, h, D // h).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert out.shape[-1] == (N + 7) // 8 * 8
    assert torch.equal(out[..., :N], ref)


@pytest.mark.parametrize("C,S,p,B", [(3, 32, 16, 2), (1, 32, 8, 2), (3, 64, 32, 3), (3, 24, 8, 3), (3, 48, 16, 3), (2, 32, 8, 1),
                                       (3, 224, 8, 2), (3, 224, 32, 3)])
@pytest.mark.parametrize("nconv", [1, 2, 3])
def test_patch_conv_tensor_core_class(ops, C, S, p, B, nconv):
    """The q/k/v convs of the tf32 / bf16 modes: implicit GEMMs on TF32 warp MMAs (tf32=True) against F.conv2d in fp64
    (operands rounded to 10 mantissa bits, fp32 accumulation: 27 / 81 products per output)."""
    img = _rand(B, C, S, S)
    x = O.patchify(img, p).contiguous()
    N, D = x.shape[1], x.shape[2]
    ws = [_rand(C, C, 3, 3, seed=10 + k, scale=0.5) for k in range(nconv)]
    xr = x.double().clone().requires_grad_(True)
    wr = [w.double().clone().requires_grad_(True) for w in ws]
    ys = [F.conv2d(xr.reshape(B * N, C, p, p), w, padding=1).reshape(B, N, D) for w in wr]
    outs = [torch.full((B, N, D), 7.0, device="cuda") for _ in range(nconv)]
    wl = [w.cuda() for w in ws]
    ops.conv3x3_fwd(x.cuda(), p, wl if nconv > 1 else wl[0], None, outs, p, p, B, C, S, S, tf32=True)
    for o, y in zip(outs, ys):
        _close(o, y, rtol=2e-3, atol_rel=2e-3, name="conv fwd (tf32 class)")
    dys = [_rand(B, N, D, seed=20 + k) for k in range(nconv)]
    sum((y * d.double()).sum() for y, d in zip(ys, dys)).backward()
    dx = torch.full((B, N, D), 7.0, device="cuda")
    ops.conv3x3_bwd_data([d.cuda() for d in dys], p, wl if nconv > 1 else wl[0], dx, p, p, B, C, S, S, tf32=True)
    _close(dx, xr.grad, rtol=2e-3, atol_rel=2e-3, name="conv dx (tf32 class)")
    base = _rand(B, N, D, seed=31)
    dx2 = base.clone().cuda()
    ops.conv3x3_bwd_data([d.cuda() for d in dys], p, wl if nconv > 1 else wl[0], dx2, p, p, B, C, S, S, accumulate=True, tf32=True)
    _close(dx2, xr.grad + base.double(), rtol=2e-3, atol_rel=2e-3, name="conv dx accumulate (tf32 class)")
    dws = [torch.zeros(C, C, 3, 3, device="cuda") for _ in range(nconv)]
    ops.conv3x3_bwd_weight(x.cuda(), p, [d.cuda() for d in dys], p, dws if nconv > 1 else dws[0], None, p, B, C, S, S, tf32=True)
    for dw, w in zip(dws, wr):
        _close(dw, w.grad, rtol=2e-3, atol_rel=2e-3, name="conv dw (tf32 class)")


@pytest.mark.parametrize("C,S,p", [(3, 32, 16), (1, 64, 32)])
def test_image_conv_on_token_layout(ops, C, S, p):
    """Reconstruction head: un-patch + 3x3 'same' conv with bias (model.py:425-428), reading tokens directly."""
    B = 2
    img = _rand(B, C, S, S).requires_grad_(True)
    w = _rand(C, C, 3, 3, seed=3, scale=0.5).requires_grad_(True)
    b = _rand(C, seed=4).requires_grad_(True)
    y = F.conv2d(img, w, b, padding=1)
    tok = O.patchify(img.detach(), p).contiguous()
    out = torch.empty(B, C, S, S, device="cuda")
    ops.conv3x3_fwd(tok.cuda(), p, w.detach().cuda(), b.detach().cuda(), [out], 0, 0, B, C, S, S)
    _close(out, y, name="head conv")
    dy = _rand(B, C, S, S, seed=5)
    (y * dy).sum().backward()
    dtok = torch.empty_like(tok, device="cuda")
    ops.conv3x3_bwd_data([dy.cuda()], 0, w.detach().cuda(), dtok, p, 0, B, C, S, S)
    _close(dtok, O.patchify(img.grad, p), name="head conv dx")
    dw, db = torch.zeros(C * C * 9, device="cuda"), torch.zeros(C, device="cuda")
    ops.conv3x3_bwd_weight(tok.cuda(), p, [dy.cuda()], 0, dw, db, 0, B, C, S, S)
    _close(dw, w.grad.reshape(-1), rtol=1e-4, name="head conv dw")
    _close(db, b.grad, rtol=1e-4, name="head conv db")


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(98, 3072, 3072), (392, 768, 768), (1568, 192, 192), (1568, 32, 192),
                                    (130, 70, 50), (64, 24, 49), (49, 49, 384), (257, 129, 1030)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_plain(ops, M, N, K, ta, tb):
    if M * N * K > 2e8 and (ta or not tb):
        pytest.skip("large shape only in the nn.Linear orientation")
    A = _rand(K, M, seed=1) if ta else _rand(M, K, seed=1)
    Bm = _rand(N, K, seed=2) if tb else _rand(K, N, seed=2)
    exp = (A.t() if ta else A).double() @ (Bm.t() if tb else Bm).double()
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A.cuda(), Bm.cuda(), out, M, N, K, trans_a=ta, trans_b=tb, lda=A.shape[1], ldb=Bm.shape[1], ldc=N)
    _close(out, exp, rtol=1e-5 * math.sqrt(K) / 4, name="gemm")


def test_gemm_epilogues_and_batch(ops):
    M, N, K = 100, 72, 40
    A, W, bias, R = _rand(M, K, seed=1), _rand(N, K, seed=2), _rand(N, seed=3), _rand(M, N, seed=4)
    pre = A @ W.t() * 0.5 + bias
    out, aux = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), out, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, bias=bias.cuda(),
             residual=R.cuda(), alpha=0.5, act=ops.ACT_GELU, aux_out=aux)
    _close(aux, pre, name="aux")
    _close(out, F.gelu(pre) + R, name="gelu+res")
    # gelu backward epilogue
    pr = pre.clone().requires_grad_(True)
    F.gelu(pr).backward(torch.ones_like(pr))
    ops.gemm(A.cuda(), W.cuda(), out, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, act=ops.ACT_GELU_BWD, aux_in=aux)
    _close(out, (A @ W.t()) * pr.grad, name="gelu bwd")
    # split-K accumulate (weight-gradient shape)
    dY, X = _rand(5000, 24, seed=5), _rand(5000, 56, seed=6)
    dW = torch.ones(24, 56, device="cuda")
    ops.gemm(dY.cuda(), X.cuda(), dW, 24, 56, 5000, trans_a=True, lda=24, ldb=56, ldc=56, accumulate=True, split_k=7)
    _close(dW, dY.double().t() @ X.double() + 1, rtol=1e-4, name="splitk")
    # head-strided batch: S[b,h] = q[b,:,h,:] k[b,:,h,:]^T
    B, h, Nt, hd = 3, 4, 49, 12
    D, ld = h * hd, 52
    q, k = _rand(B, Nt, D, seed=7), _rand(B, Nt, D, seed=8)
    S = torch.zeros(B, h, Nt, ld, device="cuda")
    ops.gemm(q.cuda(), k.cuda(), S, Nt, Nt, hd, trans_b=True, lda=D, ldb=D, ldc=ld, batch_outer=B, batch_inner=h,
             sA=(Nt * D, hd), sB=(Nt * D, hd), sC=(h * Nt * ld, Nt * ld))
    exp = torch.einsum("bihe,bjhe->bhij", q.reshape(B, Nt, h, hd), k.reshape(B, Nt, h, hd))
    _close(S[..., :Nt], exp, name="batched qk")
    assert torch.all(S[..., Nt:] == 0)


def test_gemm_dropout_matches_standalone(ops):
    M, N, K = 64, 48, 32
    A, W = _rand(M, K, seed=1), _rand(N, K, seed=2)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A.cuda(), W.cuda(), out, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, drop_p=0.3, drop_seed=1234,
             drop_stream=5)
    plain = (A @ W.t()).cuda()
    exp = ops.dropout(plain, torch.empty_like(plain), 0.3, 1234, 5)
    _close(out, exp, name="dropout epilogue")
    frac = (exp == 0).float().mean().item()
    assert 0.25 < frac < 0.35
    assert torch.allclose(exp[exp != 0], plain[exp != 0] / 0.7, rtol=1e-4)       # p is quantised to 16 bits


def test_colsum(ops):
    X = _rand(3001, 70, seed=3)
    out = torch.ones(70, device="cuda")
    ops.colsum(X.cuda(), 3001, 70, 70, out, accumulate=True)
    _close(out, X.double().sum(0) + 1, rtol=1e-5, name="colsum")


# ------------------------------------------------------------------------------------------------ re-attention
def _reattn_ref(S, scale, W, b, gamma, beta, rm, rv, train, mask=None, p=0.0):
    Pm = F.softmax(S * scale, dim=-1)
    Pd = Pm if mask is None else Pm * mask / (1 - p)
    h = S.shape[1]
    M = F.conv2d(Pd, W.reshape(h, h, 1, 1), b)
    A = F.batch_norm(M, rm, rv, gamma, beta, training=train, momentum=0.1, eps=1e-5)
    return Pm, A


@pytest.mark.parametrize("h,N", [(8, 49), (4, 196), (2, 30), (8, 70)])
@pytest.mark.parametrize("train", [False, True])
def test_reattn_forward_backward(ops, h, N, train):
    B, ld = 3, (N + 3) // 4 * 4
    scale = 0.37
    S = _rand(B, h, N, N, seed=1, scale=3.0).requires_grad_(True)
    W = _rand(h, h, seed=2, scale=0.6).requires_grad_(True)
    b = _rand(h, seed=3, scale=0.01).requires_grad_(True)
    gamma = (1 + _rand(h, seed=4, scale=0.3)).requires_grad_(True)
    beta = _rand(h, seed=5, scale=0.01).requires_grad_(True)
    rm, rv = _rand(h, seed=6, scale=0.01), (1 + _rand(h, seed=7, scale=0.3)) * 1e-3
    rm_ref, rv_ref = rm.clone(), rv.clone()
    Pm_ref, A_ref = _reattn_ref(S, scale, W, b, gamma, beta, rm_ref, rv_ref, train)
    dA = _rand(B, h, N, N, seed=8)
    (A_ref * dA).sum().backward()

    Sd = torch.zeros(B, h, N, ld, device="cuda"); Sd[..., :N] = S.detach().cuda()
    ops.softmax_rows(Sd, B * h * N, N, ld, scale)
    _close(Sd[..., :N], Pm_ref, name="softmax")
    Wd, bd, gd, btd = W.detach().cuda(), b.detach().cuda(), gamma.detach().cuda(), beta.detach().cuda()
    rmd, rvd = rm.clone().cuda(), rv.clone().cuda()
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    sums = torch.zeros(h + h * h, dtype=torch.float64, device="cuda") if train else None
    if train:
        ops.reattn_stats(Sd, B, h, N, ld, 0.0, 0, 0, sums)
    fold, saved = torch.empty(h * h + h, device="cuda"), torch.empty(2 * h, device="cuda")
    ops.reattn_bn_finalize(sums, B * N * N, h, N, Wd, bd, gd, btd, rmd, rvd, nbt, 1e-5, 0.1, train, fold, saved)
    A = torch.empty_like(Sd)
    ops.reattn_mix(Sd, A, fold, B, h, N, ld, 0.0, 0, 0)
    _close(A[..., :N], A_ref, rtol=2e-4, name="mixed map")
    assert torch.all(A[..., N:] == 0)
    if train:
        _close(rmd, rm_ref, rtol=1e-5, name="running_mean")
        _close(rvd, rv_ref, rtol=1e-4, name="running_var")
        assert nbt.item() == 1
    # backward
    dAd = torch.zeros(B, h, N, ld, device="cuda"); dAd[..., :N] = dA.cuda()
    red = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.reattn_bwd_reduce(Sd, dAd, B, h, N, ld, 0.0, 0, 0, red)
    dW, dbc, dg, dbt = (torch.zeros(h * h, device="cuda"), torch.zeros(h, device="cuda"),
                        torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"))
    coef = torch.empty(2 * h, device="cuda")
    ops.reattn_bwd_params(red, sums, B, h, N, Wd, bd, gd, saved, train, coef, dW, dbc, dg, dbt)
    ops.reattn_bwd_rows(Sd, dAd, B, h, N, ld, Wd, bd, gd, saved, coef, train, scale, 0.0, 0, 0)
    # dAd now holds dL/d(raw scores)
    _close(dAd[..., :N], S.grad, rtol=5e-4, name="dS")
    _close(dW.reshape(h, h), W.grad, rtol=5e-4, name="dW mix")
    _close(dg, gamma.grad, rtol=5e-4, name="dgamma")
    _close(dbt, beta.grad, rtol=5e-4, name="dbeta")
    if not train:
        _close(dbc, b.grad, rtol=5e-4, name="dbias mix")
    else:   # BN removes the mean: the true gradient is 0; ours must be tiny relative to |dM| mass
        assert dbc.abs().max().item() <= 1e-3 * max(1.0, W.grad.abs().max().item() * N)


def test_reattn_dropout_consistency(ops):
    """The Philox mask is identical in stats / mix / backward: compare against torch with the mask read back."""
    B, h, N, p = 2, 4, 33, 0.2
    ld = 36
    scale = 0.5
    S = _rand(B, h, N, N, seed=1, scale=2.0).requires_grad_(True)
    W = _rand(h, h, seed=2, scale=0.6).requires_grad_(True)
    b = _rand(h, seed=3, scale=0.01).requires_grad_(True)
    gamma = (1 + _rand(h, seed=4, scale=0.3)).requires_grad_(True)
    beta = _rand(h, seed=5, scale=0.01).requires_grad_(True)
    ones = torch.ones(B, h, N, ld, device="cuda")
    keep = ops.dropout(ones, torch.empty_like(ones), p, 99, 6)          # same (seed, stream, flat index) keying
    mask = (keep[..., :N] != 0).float().cpu()
    assert 0.7 < mask.mean().item() < 0.9
    Pm_ref, A_ref = _reattn_ref(S, scale, W, b, gamma, beta, torch.zeros(h), torch.ones(h), True, mask, p)
    dA = _rand(B, h, N, N, seed=8)
    (A_ref * dA).sum().backward()
    Sd = torch.zeros(B, h, N, ld, device="cuda"); Sd[..., :N] = S.detach().cuda()
    ops.softmax_rows(Sd, B * h * N, N, ld, scale)
    Wd, bd, gd, btd = W.detach().cuda(), b.detach().cuda(), gamma.detach().cuda(), beta.detach().cuda()
    sums = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.reattn_stats(Sd, B, h, N, ld, p, 99, 6, sums)
    fold, saved = torch.empty(h * h + h, device="cuda"), torch.empty(2 * h, device="cuda")
    ops.reattn_bn_finalize(sums, B * N * N, h, N, Wd, bd, gd, btd, torch.zeros(h, device="cuda"),
                           torch.ones(h, device="cuda"), None, 1e-5, 0.1, True, fold, saved)
    A = torch.empty_like(Sd)
    ops.reattn_mix(Sd, A, fold, B, h, N, ld, p, 99, 6)
    _close(A[..., :N], A_ref, rtol=2e-4, name="mixed map (dropout)")
    dAd = torch.zeros(B, h, N, ld, device="cuda"); dAd[..., :N] = dA.cuda()
    red = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.reattn_bwd_reduce(Sd, dAd, B, h, N, ld, p, 99, 6, red)
    dW, dbc, dg, dbt = (torch.zeros(h * h, device="cuda"), torch.zeros(h, device="cuda"),
                        torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"))
    coef = torch.empty(2 * h, device="cuda")
    ops.reattn_bwd_params(red, sums, B, h, N, Wd, bd, gd, saved, True, coef, dW, dbc, dg, dbt)
    ops.reattn_bwd_rows(Sd, dAd, B, h, N, ld, Wd, bd, gd, saved, coef, True, scale, p, 99, 6)
    _close(dg, gamma.grad, rtol=5e-4, name="dgamma (dropout)")
    _close(dbt, beta.grad, rtol=5e-4, name="dbeta (dropout)")
    _close(dAd[..., :N], S.grad, rtol=5e-4, name="dS (dropout)")
    _close(dW.reshape(h, h), W.grad, rtol=5e-4, name="dW mix (dropout)")


# ------------------------------------------------------------------------------------------------ layer norm
@pytest.mark.parametrize("B,N,D", [(3, 49, 3072), (2, 16, 48), (5, 7, 9)])
def test_layernorm(ops, B, N, D):
    x = (_rand(B, N, D, seed=1) * 2 + 0.3).requires_grad_(True)
    w = (1 + _rand(N, D, seed=2, scale=0.2)).requires_grad_(True)
    b = _rand(N, D, seed=3, scale=0.1).requires_grad_(True)
    y = F.layer_norm(x, (N, D), w, b, 1e-5)
    g = _rand(B, N, D, seed=4)
    (y * g).sum().backward()
    n = N * D
    stats = torch.empty(B, 2, device="cuda")
    xd = x.detach().cuda()
    ops.ln_stats(xd, B, n, 1e-5, stats)
    out = torch.empty_like(xd)
    ops.ln_apply(xd, stats, w.detach().cuda(), b.detach().cuda(), out, B, n)
    _close(out, y, name="ln fwd")
    dx, dw, db = torch.empty_like(xd), torch.zeros(N, D, device="cuda"), torch.zeros(N, D, device="cuda")
    ops.ln_bwd(g.cuda(), xd, stats, w.detach().cuda(), dx, dw, db, torch.empty(B, ops.LN_SCRATCH, device="cuda"), B, n)
    _close(dx, x.grad, rtol=1e-4, name="ln dx")
    _close(dw, w.grad, rtol=1e-5, name="ln dw")
    _close(db, b.grad, rtol=1e-5, name="ln db")


# ------------------------------------------------------------------------------------------------ losses / optimizer
@pytest.mark.parametrize("kind", ["l1", "mse", "dice"])
def test_losses(kind):
    import vit_unet_b200 as vu
    pred = _rand(4, 3, 32, 32, seed=1).requires_grad_(True)
    tgt = (_rand(4, 3, 32, 32, seed=2) > 0).float() if kind == "dice" else _rand(4, 3, 32, 32, seed=2)
    ref = {"l1": F.l1_loss, "mse": F.mse_loss, "dice": O.dice_loss}[kind](pred, tgt)
    (ref * 1.7).backward()
    pc = pred.detach().cuda().requires_grad_(True)
    got = {"l1": vu.l1_loss, "mse": vu.mse_loss, "dice": vu.dice_loss}[kind](pc, tgt.cuda())
    (got * 1.7).backward()
    _close(got, ref, name="loss")
    _close(pc.grad, pred.grad, name="dloss")


def test_adamw(ops):
    p0, g = _rand(1000, seed=1), _rand(1000, seed=2)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-2, weight_decay=0.05)
    pd, m, v = p0.clone().cuda(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    for step in range(1, 4):
        pr.grad = g * step
        opt.step()
        ops.adamw(pd, (g * step).cuda(), m, v, 1e-2, 0.9, 0.999, 1e-8, 0.05, step)
    _close(pd, pr, rtol=1e-5, name="adamw")


@pytest.mark.parametrize("h,N,p", [(8, 49, 0.0), (4, 70, 0.25), (2, 196, 0.1)])
def test_reattn_fused_passes_match_separate_kernels(ops, h, N, p):
    """vu_softmax_stats == softmax_rows + reattn_stats;  vu_reattn_mix_reduce == reattn_mix + reattn_bwd_reduce."""
    B, ld, scale = 3, (N + 3) // 4 * 4, 0.41
    S = torch.zeros(B, h, N, ld, device="cuda"); S[..., :N] = _rand(B, h, N, N, seed=1, scale=3.0).cuda()
    P1, P2 = S.clone(), S.clone()
    s1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    ops.softmax_rows(P1, B * h * N, N, ld, scale)
    ops.reattn_stats(P1, B, h, N, ld, p, 5, 2, s1)
    ops.softmax_stats(P2, B, h, N, ld, scale, p, 5, 2, s2)
    _close(P2, P1, rtol=1e-6, name="fused softmax")       # same formula; summation order differs (float4 vs scalar lanes)
    _close(s2, s1, rtol=1e-5, name="fused moments")
    fold = _rand(h * h + h, seed=2).cuda()
    dA = torch.zeros(B, h, N, ld, device="cuda"); dA[..., :N] = _rand(B, h, N, N, seed=3).cuda()
    A1, A2 = torch.empty_like(P1), torch.empty_like(P1)
    r1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); r2 = torch.zeros_like(r1)
    ops.reattn_mix(P1, A1, fold, B, h, N, ld, p, 5, 2)
    ops.reattn_bwd_reduce(P1, dA, B, h, N, ld, p, 5, 2, r1)
    ops.reattn_mix_reduce(P1, dA, A2, fold, B, h, N, ld, p, 5, 2, r2)
    assert torch.equal(A1, A2)
    _close(r2, r1, rtol=1e-6, name="fused reductions")


@pytest.mark.parametrize("N", [40, 64, 200, 328, 784])      # > 256: CTA-per-row variants
@pytest.mark.parametrize("p", [0.0, 0.2])
@pytest.mark.parametrize("train", [False, True])
def test_reattn_tensor_core_path(ops, N, p, train):
    """8 heads, no pad columns, bf16 maps: the warp-MMA formulation (vu_reattn_mma.cuh) against the CUDA-core fp32
    kernels on the same inputs.  Tolerances: TF32 rounding of centred inputs for the sums, bf16 storage for the maps."""
    B, h, ld, scale = (3 if N < 300 else 2), 8, N, 0.41
    bf = lambda t: t.to(torch.bfloat16)
    S = (_rand(B, h, N, N, seed=1, scale=3.0)).cuda()
    P1, P2 = S.clone(), S.clone()
    s1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    ops.softmax_rows(P1, B * h * N, N, ld, scale)
    ops.reattn_stats(P1, B, h, N, ld, p, 5, 2, s1)
    ops.softmax_stats(P2, B, h, N, ld, scale, p, 5, 2, s2, precision=ops.PREC_TF32)
    _close(P2, P1, rtol=1e-5, name="softmax (online, mma path)")
    _close(s2, s1, rtol=3e-4, name="moments (tf32 mma)")
    W = _rand(h, h, seed=2, scale=0.6).cuda()
    bc, gm, bt = _rand(h, seed=3, scale=0.01).cuda(), (1 + _rand(h, seed=4, scale=0.3)).cuda(), _rand(h, seed=5, scale=0.01).cuda()
    rm, rv = _rand(h, seed=6, scale=0.01).cuda(), ((1 + _rand(h, seed=7, scale=0.3)) * 1e-3).cuda()
    fold, saved = torch.empty(h * h + h, device="cuda"), torch.empty(2 * h, device="cuda")
    ops.reattn_bn_finalize(s1 if train else None, B * N * N, h, N, W, bc, gm, bt, rm, rv, None, 1e-5, 0.1, train, fold, saved)
    A1 = torch.empty_like(P1); A2 = torch.empty(B, h, N, ld, dtype=torch.bfloat16, device="cuda")
    ops.reattn_mix(P1, A1, fold, B, h, N, ld, p, 5, 2)
    ops.reattn_mix(P1, A2, fold, B, h, N, ld, p, 5, 2)
    _close(A2.float(), A1, rtol=6e-3, name="mixed map (mma, bf16)")
    dA = bf(_rand(B, h, N, N, seed=3).cuda())
    dA32 = dA.float()
    r1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); r2 = torch.zeros_like(r1)
    A3 = torch.empty_like(A2)
    ops.reattn_bwd_reduce(P1, dA32, B, h, N, ld, p, 5, 2, r1)
    ops.reattn_mix_reduce(P1, dA, A3, fold, B, h, N, ld, p, 5, 2, r2)
    assert torch.equal(A3, A2)
    r2b = torch.zeros_like(r2)                                   # A = None: reductions only (forward map kept)
    ops.reattn_mix_reduce(P1, dA, None, fold, B, h, N, ld, p, 5, 2, r2b)
    _close(r2b, r2, rtol=1e-6, name="reductions-only mode")
    _close(r2[:h], r1[:h], rtol=1e-5, name="s1 (mma)")
    _close(r2[h:], r1[h:], rtol=2e-3, name="X' (tf32 mma)")
    dW, dbc, dg, dbt = (torch.zeros(h * h, device="cuda"), torch.zeros(h, device="cuda"),
                        torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"))
    coef = torch.empty(2 * h, device="cuda")
    ops.reattn_bwd_params(r1, s1 if train else None, B, h, N, W, bc, gm, saved, train, coef, dW, dbc, dg, dbt)
    d1, d2 = dA32.clone(), dA.clone()
    ops.reattn_bwd_rows(P1, d1, B, h, N, ld, W, bc, gm, saved, coef, train, scale, p, 5, 2)
    ops.reattn_bwd_rows(P1, d2, B, h, N, ld, W, bc, gm, saved, coef, train, scale, p, 5, 2)
    _close(d2.float(), d1, rtol=1.5e-2, name="dS (mma, bf16)")
    # probabilities stored as centred bf16 (Pc = P - 1/N): same kernels, half the bytes for P
    Pc = torch.empty(B, h, N, ld, dtype=torch.bfloat16, device="cuda")
    s3 = torch.zeros_like(s1)
    S3 = S.clone()
    ops.softmax_stats(S3, B, h, N, ld, scale, p, 5, 2, s3, precision=ops.PREC_TF32, Pc=Pc)
    assert torch.equal(S3, S)                                   # scores untouched
    _close(Pc.float(), P1 - 1.0 / N, rtol=4e-3, name="centred bf16 probabilities")
    _close(s3, s1, rtol=8e-3, name="moments of the rounded map")
    A4 = torch.empty_like(A2)
    ops.reattn_mix(Pc, A4, fold, B, h, N, ld, p, 5, 2)
    _close(A4.float(), A1, rtol=1.2e-2, name="mixed map (centred bf16 P)")
    r3 = torch.zeros_like(r1); A5 = torch.empty_like(A2)
    ops.reattn_mix_reduce(Pc, dA, A5, fold, B, h, N, ld, p, 5, 2, r3)
    assert torch.equal(A5, A4)
    _close(r3[h:], r1[h:], rtol=4e-3, name="X' (centred bf16 P)")
    d3 = dA.clone()
    ops.reattn_bwd_rows(Pc, d3, B, h, N, ld, W, bc, gm, saved, coef, train, scale, p, 5, 2)
    _close(d3.float(), d1, rtol=2.5e-2, name="dS (centred bf16 P)")


@pytest.mark.parametrize("N,p,train", [(196, 0.2, True), (52, 0.0, True), (196, 0.0, False)])
def test_reattn_tensor_core_path_fp32_maps(ops, N, p, train):
    """fp32 maps with tf32=True (VU_MAP_TF32_MIX): N % 4 == 0 but not % 8 (Base level 1, N = 196) still runs the
    warp-MMA kernels; results against the exact CUDA-core kernels at TF32-class tolerance."""
    B, h, ld, scale = 3, 8, N, 0.41
    S = (_rand(B, h, N, N, seed=1, scale=3.0)).cuda()
    P1, P2 = S.clone(), S.clone()
    s1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    ops.softmax_stats(P1, B, h, N, ld, scale, p, 5, 2, s1)
    ops.softmax_stats(P2, B, h, N, ld, scale, p, 5, 2, s2, precision=ops.PREC_TF32)
    _close(P2, P1, rtol=1e-5, name="softmax")
    _close(s2, s1, rtol=3e-4, name="moments")
    W = _rand(h, h, seed=2, scale=0.6).cuda()
    bc, gm, bt = _rand(h, seed=3, scale=0.01).cuda(), (1 + _rand(h, seed=4, scale=0.3)).cuda(), _rand(h, seed=5, scale=0.01).cuda()
    rm, rv = _rand(h, seed=6, scale=0.01).cuda(), ((1 + _rand(h, seed=7, scale=0.3)) * 1e-3).cuda()
    fold, saved = torch.empty(h * h + h, device="cuda"), torch.empty(2 * h, device="cuda")
    ops.reattn_bn_finalize(s1 if train else None, B * N * N, h, N, W, bc, gm, bt, rm, rv, None, 1e-5, 0.1, train, fold, saved)
    A1, A2 = torch.empty_like(P1), torch.empty_like(P1)
    ops.reattn_mix(P1, A1, fold, B, h, N, ld, p, 5, 2)
    ops.reattn_mix(P1, A2, fold, B, h, N, ld, p, 5, 2, tf32=True)
    _close(A2, A1, rtol=2e-3, name="mixed map (tf32 mix)")
    assert not torch.equal(A1, A2)                       # really the tensor-core kernel
    dA = _rand(B, h, N, N, seed=3).cuda()
    r1 = torch.zeros(h + h * h, dtype=torch.float64, device="cuda"); r2 = torch.zeros_like(r1)
    A3, A4 = torch.empty_like(P1), torch.empty_like(P1)
    ops.reattn_mix_reduce(P1, dA, A3, fold, B, h, N, ld, p, 5, 2, r1)
    ops.reattn_mix_reduce(P1, dA, A4, fold, B, h, N, ld, p, 5, 2, r2, tf32=True)
    assert torch.equal(A4, A2)
    _close(r2, r1, rtol=2e-3, name="reductions (tf32 mma)")
    dW, dbc, dg, dbt = (torch.zeros(h * h, device="cuda"), torch.zeros(h, device="cuda"),
                        torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"))
    coef = torch.empty(2 * h, device="cuda")
    ops.reattn_bwd_params(r1, s1 if train else None, B, h, N, W, bc, gm, saved, train, coef, dW, dbc, dg, dbt)
    d1, d2 = dA.clone(), dA.clone()
    ops.reattn_bwd_rows(P1, d1, B, h, N, ld, W, bc, gm, saved, coef, train, scale, p, 5, 2)
    ops.reattn_bwd_rows(P1, d2, B, h, N, ld, W, bc, gm, saved, coef, train, scale, p, 5, 2, tf32=True)
    _close(d2, d1, rtol=1.2e-2, name="dS (tf32 mix)")


def test_psnr_and_input_pipeline(ops):
    """N2 / N4: device PSNR vs the skimage formula; uint8 HWC -> normalised float CHW vs numpy."""
    g = torch.Generator().manual_seed(0)
    y = torch.rand(5, 3, 40, 40, generator=g)
    y[1] -= 0.5                                       # negative values -> skimage data_range 2 for that image
    out = y + 0.05 * torch.randn(5, 3, 40, 40, generator=g)
    got = ops.psnr(out.cuda(), y.cuda()).cpu()
    mse = ((out - y) ** 2).flatten(1).mean(1).double()
    dr = torch.tensor([1.0, 2.0, 1.0, 1.0, 1.0], dtype=torch.float64)
    _close(got, 10 * torch.log10(dr * dr / mse), rtol=1e-5, name="psnr")
    _close(ops.psnr(out.cuda(), y.cuda(), 4.0), 10 * torch.log10(16.0 / mse), rtol=1e-5, name="psnr fixed range")
    img = torch.randint(0, 256, (3, 20, 24, 3), dtype=torch.uint8, generator=g)
    exp = ((img.float() / 255.0 - 0.456) / 0.224).permute(0, 3, 1, 2)
    _close(ops.u8hwc_to_chw(img.cuda(), 1 / 255.0, 0.456, 0.224), exp, rtol=1e-5, name="u8hwc->chw")


def test_dice_global_batch_semantics_from_partial_sums(ops):
    """Data-parallel soft-Dice (README.md:96-101 is a whole-batch ratio): adding the per-shard sums and finalising once
    equals the Dice of the whole batch, and the shard gradients computed from the GLOBAL sums are the rows of the
    whole-batch gradient (what losses.dice_loss(group=...) does with a 3-double all-reduce)."""
    g = torch.Generator().manual_seed(4)
    pred, tgt = torch.rand(4, 1, 32, 32, generator=g).cuda(), (torch.rand(4, 1, 32, 32, generator=g) > 0.6).float().cuda()
    full = torch.empty(4, dtype=torch.float64, device="cuda"); lf = torch.empty((), device="cuda")
    ops.loss_fwd("dice", pred, tgt, full, lf)
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    for a, b in ((pred[:2].contiguous(), tgt[:2].contiguous()), (pred[2:].contiguous(), tgt[2:].contiguous())):
        s = torch.empty(4, dtype=torch.float64, device="cuda"); l = torch.empty((), device="cuda")
        ops.loss_fwd("dice", a, b, s, l)
        sums += s
    lg = torch.empty((), device="cuda")
    ops.loss_finalize("dice", pred.numel(), sums, lg)
    assert abs(lg.item() - lf.item()) <= 1e-6
    exp = 1 - (2 * (pred.double() * tgt.double()).sum() + 1) / (pred.double().sum() + tgt.double().sum() + 1)
    assert abs(lg.item() - exp.item()) <= 1e-6
    one = torch.ones(1, device="cuda")
    dfull = torch.empty_like(pred); ops.loss_bwd("dice", pred, tgt, full, one, dfull)
    dpart = torch.empty(2, 1, 32, 32, device="cuda")
    ops.loss_bwd("dice", pred[2:].contiguous(), tgt[2:].contiguous(), sums, one, dpart)
    assert torch.allclose(dpart, dfull[2:], rtol=1e-6, atol=1e-9)
