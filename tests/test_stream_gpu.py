"""Streamed Re-Attention forward (vu_reattn_stream.cu) against (a) the math of the reference written with torch ops
in fp64 (model.py:155-161: softmax -> [dropout] -> 1x1 conv + BatchNorm folded to an h x h affine -> @ v) and
(b) the materialised kernel chain with the SAME dropout seed (the masks agree element for element).
Precision class: TF32 scores, bf16 A.V operands -> 1e-2 of the output scale (north_star), observed ~2e-3."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(8, 24, 784), (8, 24, 128), (8, 8, 784), (8, 8, 208), (4, 12, 784), (4, 48, 336), (4, 12, 64)]


@pytest.fixture(scope="module")
def ops():
    from vit_unet_b200 import ops as _ops
    return _ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(*shape, generator=g) * 2 - 1) * scale).cuda()


def _inputs(B, h, hd, N, seed=0):
    D = h * hd
    q, k, v = _rand(B, N, D, seed=seed + 1, scale=1.5), _rand(B, N, D, seed=seed + 2, scale=1.5), _rand(B, N, D, seed=seed + 3)
    fold = torch.cat([_rand(h, h, seed=seed + 4).reshape(-1) * 40.0, _rand(h, seed=seed + 5) * 0.05])
    return q, k, v, fold.contiguous()


def _heads(t, h):                                  # (B, N, D) -> (B, h, N, hd), fp64
    B, N, D = t.shape
    return t.double().reshape(B, N, h, D // h).permute(0, 2, 1, 3)


def _ref_probs(q, k, h):
    hd = q.shape[-1] // h
    return torch.softmax(_heads(q, h) @ _heads(k, h).transpose(-1, -2) * hd ** -0.5, dim=-1)


def _ref_out(Pd, v, fold, h):
    alpha, beta = fold[:h * h].double().reshape(h, h), fold[h * h:].double()
    A = torch.einsum("hg,bgij->bhij", alpha, Pd) + beta.view(1, h, 1, 1)
    O = A @ _heads(v, h)                           # (B, h, N, hd)
    B, _, N, hd = O.shape
    return O.permute(0, 2, 1, 3).reshape(B, N, h * hd)


def _close(a, b, tol, name):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert err <= tol, f"{name}: {err:.3e} > {tol}"


@pytest.mark.parametrize("h,hd,N", SHAPES)
def test_stream_eval_forward(ops, h, hd, N):
    B = 3
    assert ops.reattn_stream_supported(h, hd, N)
    q, k, v, fold = _inputs(B, h, hd, N)
    vt = ops.heads_transpose_bf16(v, B, N, h * hd, h)
    O = torch.full((B, N, h * hd), 7.0, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_EVAL, q, k, vt, O, fold, None, None, None, B, h, N, hd, hd ** -0.5)
    _close(O, _ref_out(_ref_probs(q, k, h), v, fold, h), 1e-2, "eval O")


@pytest.mark.parametrize("h,hd,N", SHAPES)
def test_stream_train_statistics_and_apply_no_dropout(ops, h, hd, N):
    B = 2
    q, k, v, fold = _inputs(B, h, hd, N, seed=10)
    rowc = torch.empty(B, h, N, device="cuda")
    sums = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    pc = torch.empty(B, h, N, N, dtype=torch.bfloat16, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_STATS, q, k, None, None, None, rowc, sums, pc, B, h, N, hd, hd ** -0.5)
    P = _ref_probs(q, k, h)
    S2 = (_heads(q, h) @ _heads(k, h).transpose(-1, -2)) * (hd ** -0.5 / math.log(2.0))
    _close(rowc, torch.logsumexp(S2 * math.log(2.0), dim=-1) / math.log(2.0), 2e-3, "row constants")
    Pc = P - 1.0 / N
    _close(pc.float(), Pc, 1e-2, "centred bf16 probabilities")
    assert sums[:h].abs().max().item() <= 1e-3 * B * N          # rows sum to 1: centred sums vanish up to round-off
    G = torch.einsum("bgij,bhij->gh", Pc, Pc)
    _close(sums[h:].reshape(h, h), G, 5e-3, "G'")
    vt = ops.heads_transpose_bf16(v, B, N, h * hd, h)
    O = torch.empty(B, N, h * hd, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_APPLY, q, k, vt, O, fold, rowc, None, None, B, h, N, hd, hd ** -0.5)
    _close(O, _ref_out(P, v, fold, h), 1e-2, "apply O")


@pytest.mark.parametrize("h,hd,N", [(8, 24, 784), (8, 24, 128), (8, 8, 208)])
def test_stream_dropout_matches_materialised_chain(ops, h, hd, N):
    """Same seed / stream id => the streamed kernels and the materialised kernels drop the same elements: moments and
    outputs agree to the precision class, and the masks can be read back exactly."""
    B, p, seed, sid = 2, 0.25, 1234, 5
    D, scale = h * hd, hd ** -0.5
    q, k, v, fold = _inputs(B, h, hd, N, seed=20)
    rowc = torch.empty(B, h, N, device="cuda")
    sums = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_STATS, q, k, None, None, None, rowc, sums, None, B, h, N, hd, scale, p, seed, sid)
    # materialised chain (fp32 exact kernels)
    S = torch.empty(B, h, N, N, device="cuda")
    ops.gemm(q, k, S, N, N, hd, trans_b=True, lda=D, ldb=D, ldc=N, batch_outer=B, batch_inner=h, sA=(N * D, hd),
             sB=(N * D, hd), sC=(h * N * N, N * N), precision=ops.PREC_FP32)
    sums_m = torch.zeros_like(sums)
    ops.softmax_stats(S, B, h, N, N, scale, p, seed, sid, sums_m, precision=ops.PREC_FP32)      # S -> P in place
    _close(sums[:h], sums_m[:h], 2e-2, "s' with dropout")
    _close(sums[h:], sums_m[h:], 1e-2, "G' with dropout")
    A = torch.empty_like(S)
    ops.reattn_mix(S, A, fold, B, h, N, N, p, seed, sid)
    O_m = (A.double() @ _heads(v, h)).permute(0, 2, 1, 3).reshape(B, N, D)
    vt = ops.heads_transpose_bf16(v, B, N, D, h)
    O = torch.empty(B, N, D, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_APPLY, q, k, vt, O, fold, rowc, None, None, B, h, N, hd, scale, p, seed, sid)
    _close(O, O_m, 1e-2, "apply O with dropout")
    # cached keep-bits (written by the statistics launch, read by the apply launch) give the same result as re-hashing
    mask = torch.zeros(ops.stream_mask_bytes(B, N), dtype=torch.uint8, device="cuda")
    sums_c = torch.zeros_like(sums)
    ops.reattn_stream_fwd(ops.STREAM_STATS, q, k, None, None, None, rowc, sums_c, None, B, h, N, hd, scale, p, seed, sid, mask=mask)
    Oc = torch.empty_like(O)
    Amap = torch.empty(B, h, N, N, dtype=torch.bfloat16, device="cuda")
    ops.reattn_stream_fwd(ops.STREAM_APPLY, q, k, vt, Oc, fold, rowc, None, None, B, h, N, hd, scale, p, seed, sid, mask=mask,
                          amap=Amap)
    assert torch.equal(Oc, O)
    _close(Amap.float(), A, 1e-2, "mixed map written for the backward pass")
    kept = sum(bin(int(b)).count("1") for b in mask[:4096].cpu().tolist()) / (4096 * 8) * (8 // h)   # 4 heads: low half used
    assert abs(kept - (1 - p)) < 0.02, kept
    # a different seed must give a different result (the mask is really applied)
    O2 = torch.empty_like(O)
    ops.reattn_stream_fwd(ops.STREAM_APPLY, q, k, vt, O2, fold, rowc, None, None, B, h, N, hd, scale, p, seed + 1, sid)
    assert (O2 - O).abs().max().item() > 1e-3 * O.abs().max().item()


def test_stream_refuses_unsupported_shapes(ops):
    from vit_unet_b200._lib import VuError
    assert not ops.reattn_stream_supported(8, 96, 196) and not ops.reattn_stream_supported(8, 24, 100)
    q = torch.zeros(1, 100, 192, device="cuda")
    with pytest.raises(VuError):
        ops.reattn_stream_fwd(ops.STREAM_EVAL, q, q, torch.zeros(1, 8, 24, 104, dtype=torch.bfloat16, device="cuda"),
                              torch.zeros_like(q), torch.zeros(72, device="cuda"), None, None, None, 1, 8, 100, 24, 0.2)


# ------------------------------------------------------------------------------------------------ streamed backward
def _bn_params(h, seed):
    W = _rand(h, h, seed=seed + 1) / h ** 0.5
    bconv = _rand(h, seed=seed + 2) * 0.05
    gamma = 1.0 + 0.3 * _rand(h, seed=seed + 3)
    beta = 0.002 * _rand(h, seed=seed + 4)
    rmean = 0.01 * _rand(h, seed=seed + 5) + 1.0 / 784
    rvar = (1e-6 * (1.0 + 0.25 * _rand(h, seed=seed + 6))).abs()
    return [t.contiguous() for t in (W, bconv, gamma, beta, rmean, rvar)]


@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("p", [0.0, 0.25])
@pytest.mark.parametrize("h,hd,N", [(8, 24, 784), (8, 24, 128), (4, 12, 336), (4, 48, 336), (8, 8, 208), (4, 12, 3136)])
def test_stream_backward_matches_materialised_chain(ops, h, hd, N, p, train):
    """vu_reattn_stream_bwd_reduce / _bwd_ds against the exact fp32 materialised kernels (scores -> softmax_stats ->
    bn_finalize -> dA = dO v^T -> bwd_reduce -> bwd_params -> bwd_rows -> dq = dS k) on the same inputs and dropout seed."""
    if p > 0 and not train:
        pytest.skip("dropout is a train-mode feature")
    B, seed, sid = (1 if N > 1024 else 2), 99, 3
    D, scale = h * hd, hd ** -0.5
    q, k, v, _ = _inputs(B, h, hd, N, seed=30)
    dO = _rand(B, N, D, seed=40)
    W, bconv, gamma, beta, rmean, rvar = _bn_params(h, 50)
    # ---- materialised reference chain (fp32 kernels)
    S = torch.empty(B, h, N, N, device="cuda")
    ops.gemm(q, k, S, N, N, hd, trans_b=True, lda=D, ldb=D, ldc=N, batch_outer=B, batch_inner=h, sA=(N * D, hd),
             sB=(N * D, hd), sC=(h * N * N, N * N), precision=ops.PREC_FP32)
    sums_m = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.softmax_stats(S, B, h, N, N, scale, p, seed, sid, sums_m, precision=ops.PREC_FP32)       # S -> P
    fold_m, bn_m = torch.empty(h * h + h, device="cuda"), torch.empty(2 * h, device="cuda")
    ops.reattn_bn_finalize(sums_m if train else None, B * N * N, h, N, W, bconv, gamma, beta, rmean.clone(), rvar.clone(),
                           None, 1e-5, 0.1, train, fold_m, bn_m)
    dA = torch.empty(B, h, N, N, device="cuda")
    ops.gemm(dO, v, dA, N, N, hd, trans_b=True, lda=D, ldb=D, ldc=N, batch_outer=B, batch_inner=h, sA=(N * D, hd),
             sB=(N * D, hd), sC=(h * N * N, N * N), precision=ops.PREC_FP32)
    red_m = torch.zeros(h + h * h, dtype=torch.float64, device="cuda")
    ops.reattn_bwd_reduce(S, dA, B, h, N, N, p, seed, sid, red_m)
    coef_m = torch.empty(2 * h, device="cuda")
    dW, db, dg, dbt = (torch.zeros(h * h, device="cuda"), torch.zeros(h, device="cuda"), torch.zeros(h, device="cuda"),
                       torch.zeros(h, device="cuda"))
    ops.reattn_bwd_params(red_m, sums_m if train else None, B, h, N, W, bconv, gamma, bn_m, train, coef_m, dW, db, dg, dbt)
    ops.reattn_bwd_rows(S, dA, B, h, N, N, W, bconv, gamma, bn_m, coef_m, train, scale, p, seed, sid)     # dA -> dS in place
    dq_m = torch.einsum("bhij,bhje->bihe", dA.double(), _heads(k, h)).reshape(B, N, D)
    # ---- streamed chain
    rowc = torch.empty(B, h, N, device="cuda")
    sums = torch.zeros_like(sums_m)
    pc = torch.empty(B, h, N, N, dtype=torch.bfloat16, device="cuda")
    mask = torch.zeros(ops.stream_mask_bytes(B, N), dtype=torch.uint8, device="cuda") if p > 0 else None
    ops.reattn_stream_fwd(ops.STREAM_STATS, q, k, None, None, None, rowc, sums, pc, B, h, N, hd, scale, p, seed, sid, mask=mask)
    red = torch.zeros_like(red_m)
    ops.reattn_stream_bwd_reduce(pc, mask, dO, v, red, B, h, N, hd, p, seed, sid)
    _close(red[:h], red_m[:h], 5e-3, "s1")
    # X' = sum dA (Pd - 1/N) is a signed sum of ~B N^2 terms that largely cancel; the streamed side reads the bf16-ROUNDED
    # centred probabilities (2^-9 relative per term), the reference chain the fp32 ones: a few per cent of the largest entry
    _close(red[h:], red_m[h:], 6e-2, "X'")
    kt = ops.heads_transpose_bf16(k, B, N, D, h)
    dS = torch.empty(B, h, N, N, dtype=torch.bfloat16, device="cuda")
    dq = torch.empty(B, N, D, device="cuda")
    # same coefficients on both sides: this compares the kernels, not the amplification of the reductions' rounding
    ops.reattn_stream_bwd_ds(pc, mask, dO, v, kt, dS, dq, W, bconv, gamma, bn_m, coef_m, train, B, h, N, hd, p, seed, sid)
    _close(dS.float(), dA, 2e-2, "dS")
    _close(dq, dq_m, 2e-2, "dq")
    if p > 0:      # re-hashing instead of the cached keep-bits gives the same gradients
        dS2, dq2 = torch.empty_like(dS), torch.empty_like(dq)
        ops.reattn_stream_bwd_ds(pc, None, dO, v, kt, dS2, dq2, W, bconv, gamma, bn_m, coef_m, train, B, h, N, hd, p, seed, sid)
        assert torch.equal(dS2, dS) and torch.equal(dq2, dq)
