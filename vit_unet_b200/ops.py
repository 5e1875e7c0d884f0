"""Tensor-level wrappers over the C ABI.  PyTorch is used only for device memory and streams.

Every function validates device/dtype/contiguity, launches on ``torch.cuda.current_stream()`` and raises
``VuError`` on failure.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmDesc, VuError, call

ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2
PREC_FP32, PREC_TF32 = 0, 1
LOSS_KINDS = {"l1": 0, "mse": 1, "dice": 2}


MAP_BF16, MAP_P_CENTRED_BF16, MAP_TF32_MIX = 1, 2, 4       # include/vit_unet_b200.h VU_MAP_*


def _map(t: torch.Tensor, name: str):
    """attention map buffer: fp32 or bf16 -> (pointer, is_bf16)"""
    if t.dtype == torch.bfloat16:
        return _chk(t, name, torch.bfloat16), 1
    return _chk(t, name), 0


def _pmap(P: torch.Tensor, bf_maps: int, tf32: bool = False):
    """probabilities: fp32, or centred bf16 (only with bf16 maps) -> (pointer, map_fmt flags, bytes per element).
    tf32: the head mixing may run on TF32 warp MMAs even with fp32 maps (tensor-core precision class)."""
    pp, pbf = _map(P, "P")
    if pbf and not bf_maps:
        raise VuError("centred bf16 probabilities need bf16 mixed / gradient maps")
    fmt = (MAP_BF16 if bf_maps else 0) | (MAP_P_CENTRED_BF16 if pbf else 0) | (MAP_TF32_MIX if tf32 else 0)
    return pp, fmt, (2.0 if pbf else 4.0)


def reattn_tensor_core_path(h: int, N: int, ld: int) -> bool:
    """True where the C side runs the 8-head warp-MMA map kernels (and accepts centred bf16 probabilities)."""
    return bool(_lib.load().vu_reattn_tensor_core_path(h, N, ld))


def _chk(t: torch.Tensor, name: str, dtype=torch.float32) -> int:
    if not t.is_cuda:
        raise VuError(f"{name}: tensor must live on a CUDA device (vit_unet_b200 has no CPU path)")
    if t.dtype != dtype:
        raise VuError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise VuError(f"{name}: tensor must be contiguous")
    return t.data_ptr()


def _opt(t: Optional[torch.Tensor], name: str, dtype=torch.float32):
    return None if t is None else _chk(t, name, dtype)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


# ----------------------------------------------------------------------------------------- kernel timing
class KernelTimer:
    """CUDA-event timing of every launch on the launching stream, with the ALGORITHMIC flops / HBM bytes of the op
    (what the math needs, not what the kernel happens to move) -> live roofline table in bench.py."""

    def __init__(self):
        self.records = []          # (class name, flops, bytes, start_event, end_event)

    def summary(self, peak_tflops, peak_gbs):
        torch.cuda.synchronize()
        agg = {}
        for name, flops, nbytes, e0, e1 in self.records:
            a = agg.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            a["ms"] += e0.elapsed_time(e1); a["flops"] += flops; a["bytes"] += nbytes; a["launches"] += 1
        total = sum(a["ms"] for a in agg.values()) or 1.0
        by = {}
        for k, v in agg.items():
            sec = max(v["ms"], 1e-9) * 1e-3
            tf, gbs = v["flops"] / sec / 1e12, v["bytes"] / sec / 1e9
            by[k] = {"ms": round(v["ms"], 3), "share": round(v["ms"] / total, 4), "launches": v["launches"], "alg_bytes": v["bytes"],
                     "tflops": round(tf, 2), "gbs": round(gbs, 1), "tensor_frac": round(tf / peak_tflops, 4),
                     "hbm_frac": round(gbs / peak_gbs, 4)}
        return {"by_kernel": by, "total_kernel_ms": total}


_TIMER = {"t": None}
_SHAPE_CLASSES = __import__("os").environ.get("VU_TIMER_SHAPES", "0") == "1"


def set_kernel_timer(t):
    _TIMER["t"] = t


def _call(name, *args, flops=0.0, nbytes=0.0, cls=None):
    t = _TIMER["t"]
    if t is None:
        call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    t.records.append((cls or name, float(flops), float(nbytes), e0, e1))


# ----------------------------------------------------------------------------------------- layout
def repatch(x, out, B, Cc, H, W, p_in, p_out):
    _call("vu_repatch", _chk(x, "in"), _chk(out, "out"), B, Cc, H, W, p_in, p_out, _stream(), nbytes=2 * 4.0 * B * Cc * H * W)
    return out


def heads_transpose_bf16(x, B, N, D, h):
    """(B,N,D) fp32 -> (B,h,D/h,ceil8(N)) bf16, dst[b,h,e,n] = x[b,n,h*hd+e]"""
    ldn = (N + 7) // 8 * 8
    out = torch.empty((B, h, D // h, ldn), dtype=torch.bfloat16, device=x.device)
    _call("vu_heads_transpose_bf16", _chk(x, "src"), _chk(out, "dst", torch.bfloat16), B, N, D, h, ldn, _stream(),
          nbytes=6.0 * B * N * D)
    return out


def pe_fwd(x, p_in, table, p_table, out, p_out, B, Cc, H, W):
    _call("vu_pe_fwd", _chk(x, "in"), p_in, _chk(table, "table"), p_table, _chk(out, "out"), p_out,
          B, Cc, H, W, _stream(), nbytes=2 * 4.0 * B * Cc * H * W)
    return out


def pe_bwd_table(dout, p_out, dtable, p_table, B, Cc, H, W, accumulate=False):
    _call("vu_pe_bwd_table", _chk(dout, "dout"), p_out, _chk(dtable, "dtable"), p_table, B, Cc, H, W,
          int(accumulate), _stream(), nbytes=4.0 * B * Cc * H * W)
    return dtable


# ----------------------------------------------------------------------------------------- convs
def _filters(w, name):
    """one contiguous tensor holding the filters of all fused convs, or a list of per-conv tensors -> three pointers"""
    if isinstance(w, (list, tuple)):
        ptrs = [_chk(t, f"{name}{i}") for i, t in enumerate(w)] + [None] * (3 - len(w))
        return ptrs[0], ptrs[1], ptrs[2]
    return _chk(w, name), None, None


def conv3x3_fwd(x, p_x, w, bias, outs, p_out, border_p, B, Cc, H, W, tf32=False):
    """tf32: tensor-core precision class (the per-patch convs may run as implicit GEMMs on TF32 warp MMAs)"""
    n = len(outs)
    ptrs = [_chk(o, f"out{i}") for i, o in enumerate(outs)] + [None] * (3 - n)
    w0, w1, w2 = _filters(w, "w")
    _call("vu_conv3x3_fwd", _chk(x, "x"), p_x, w0, w1, w2, _opt(bias, "bias"), n, ptrs[0], ptrs[1], ptrs[2],
          p_out, border_p, B, Cc, H, W, int(tf32), _stream(), nbytes=(1 + n) * 4.0 * B * Cc * H * W, flops=18.0 * n * Cc * Cc * B * H * W)
    return outs


def conv3x3_bwd_data(dys, p_dy, w, dx, p_dx, border_p, B, Cc, H, W, accumulate=False, tf32=False):
    n = len(dys)
    ptrs = [_chk(o, f"dy{i}") for i, o in enumerate(dys)] + [None] * (3 - n)
    w0, w1, w2 = _filters(w, "w")
    _call("vu_conv3x3_bwd_data", ptrs[0], ptrs[1], ptrs[2], p_dy, w0, w1, w2, n, _chk(dx, "dx"), p_dx,
          border_p, B, Cc, H, W, int(accumulate), int(tf32), _stream(), nbytes=(1 + n + int(accumulate)) * 4.0 * B * Cc * H * W,
          flops=18.0 * n * Cc * Cc * B * H * W)
    return dx


def conv3x3_bwd_weight(x, p_x, dys, p_dy, dw, dbias, border_p, B, Cc, H, W, tf32=False):
    """dw: one contiguous [nconv][C][C][3][3] tensor, or a list of per-conv gradient tensors (accumulated in place)"""
    n = len(dys)
    ptrs = [_chk(o, f"dy{i}") for i, o in enumerate(dys)] + [None] * (3 - n)
    d0, d1, d2 = _filters(dw, "dw")
    _call("vu_conv3x3_bwd_weight", _chk(x, "x"), p_x, ptrs[0], ptrs[1], ptrs[2], p_dy, n, d0, d1, d2,
          _opt(dbias, "dbias"), border_p, B, Cc, H, W, int(tf32), _stream(), nbytes=(1 + n) * 4.0 * B * Cc * H * W,
          flops=18.0 * n * Cc * Cc * B * H * W)


def zeros(shape, dtype, device):
    """torch.empty + a stream-ordered cudaMemsetAsync (accumulators, gradient buffers): no fill kernel"""
    t = torch.empty(shape, dtype=dtype, device=device)
    if t.numel():
        call("vu_zero", t.data_ptr(), t.numel() * t.element_size(), _stream())
    return t


# ----------------------------------------------------------------------------------------- GEMM
def gemm(A, Bm, Cm, M, N, K, *, trans_a=False, trans_b=False, lda, ldb, ldc,
         bias=None, residual=None, ldr=0, aux_in=None, aux_out=None, ldaux=0,
         batch_outer=1, batch_inner=1, sA=(0, 0), sB=(0, 0), sC=(0, 0),
         alpha=1.0, act=ACT_NONE, accumulate=False, split_k=1,
         drop_p=0.0, drop_seed=0, drop_stream=0, precision=PREC_FP32):
    """C = act(alpha * op(A) @ op(B) + bias) [dropout] + residual, batched over (outer, inner)."""
    d = GemmDesc()
    d.a_bf16, d.b_bf16, d.c_bf16 = int(A.dtype == torch.bfloat16), int(Bm.dtype == torch.bfloat16), int(Cm.dtype == torch.bfloat16)
    d.A, d.B, d.C = _chk(A, "A", A.dtype if d.a_bf16 else torch.float32), _chk(Bm, "B", Bm.dtype if d.b_bf16 else torch.float32), \
        _chk(Cm, "C", Cm.dtype if d.c_bf16 else torch.float32)
    d.bias, d.residual = _opt(bias, "bias"), _opt(residual, "residual")
    aux = aux_in if aux_in is not None else aux_out
    d.aux_bf16 = int(aux is not None and aux.dtype == torch.bfloat16)
    adt = torch.bfloat16 if d.aux_bf16 else torch.float32
    d.aux_in, d.aux_out = _opt(aux_in, "aux_in", adt), _opt(aux_out, "aux_out", adt)
    d.M, d.N, d.K = M, N, K
    d.trans_a, d.trans_b = int(trans_a), int(trans_b)
    d.lda, d.ldb, d.ldc, d.ldr, d.ldaux = lda, ldb, ldc, ldr or ldc, ldaux or ldc
    d.batch_outer, d.batch_inner = batch_outer, batch_inner
    d.sAo, d.sAi = sA
    d.sBo, d.sBi = sB
    d.sCo, d.sCi = sC
    d.alpha, d.act, d.accumulate, d.split_k = alpha, act, int(accumulate), split_k
    d.drop_p, d.drop_seed, d.drop_stream = drop_p, drop_seed, drop_stream
    d.precision = precision
    nb = batch_outer * batch_inner
    # class names of the live roofline table: the batched map products keep their round-1 names whatever the operand type
    kind = ("gemm_tcgen05_bf16" if (d.a_bf16 and nb == 1) else "gemm_tcgen05_tf32") if precision == PREC_TF32 else "gemm_simt_fp32"
    if (precision == PREC_TF32 and not trans_a and trans_b and K <= 128 and ((N + 7) // 8 * 8) * (K + 8) * 4 <= 200 * 1024
            and M >= 64 and N >= 64 and not d.a_bf16
            and bias is None and residual is None and aux_in is None and aux_out is None and act == ACT_NONE
            and not accumulate and split_k <= 1 and drop_p == 0.0):
        kind = "gemm_mma_tf32"          # vu_gemm_scores.cu: warp-MMA write-stream kernel (same routing rule as the C side)
    # classes: contractions over the head dim that WRITE an NxN map / contractions that READ a map / token GEMMs
    if nb > 1:
        cls = kind + (":map_out(QK^T,dA)" if N == M and K < N else ":map_in(PV,dV,dQ,dK)")
    else:
        cls = kind + ":tokens(proj,FF,dgrad,wgrad)"
    if _SHAPE_CLASSES:           # VU_TIMER_SHAPES=1: one class per GEMM shape / epilogue (tools/step_gemm_shapes.py)
        cls += f" M={M} N={N} K={K} t={int(trans_a)}{int(trans_b)} ep={'b' if bias is not None else ''}{'r' if residual is not None else ''}" \
               f"{'g' if act == ACT_GELU else ('G' if act == ACT_GELU_BWD else '')}{'d' if drop_p > 0 else ''}{'a' if accumulate else ''} sk={split_k}"
    extra = (residual is not None) + int(accumulate) + (aux_in is not None) + (aux_out is not None)
    ea, eb, ec = (2.0 if d.a_bf16 else 4.0), (2.0 if d.b_bf16 else 4.0), (2.0 if d.c_bf16 else 4.0)
    _call("vu_gemm", C.byref(d), _stream(), flops=2.0 * M * N * K * nb,
          nbytes=nb * (ea * M * K + eb * K * N + ec * M * N + 4.0 * M * N * extra), cls=cls)
    return Cm


def colsum(X, M, N, ld, out, accumulate=False):
    bf = X.dtype == torch.bfloat16
    _call("vu_colsum", _chk(X, "X", X.dtype if bf else torch.float32), int(bf), M, N, ld, _chk(out, "out"), int(accumulate),
          _stream(), nbytes=(2.0 if bf else 4.0) * M * N)
    return out


# ----------------------------------------------------------------------------------------- re-attention
def softmax_rows(S, rows, N, ld, scale):
    _call("vu_softmax_rows", _chk(S, "S"), rows, N, ld, scale, _stream(), nbytes=8.0 * rows * N)


def softmax_stats(S, B, h, N, ld, scale, drop_p, seed, sid, sums, precision=PREC_FP32, Pc=None):
    """in place (Pc is None) or S -> centred bf16 Pc"""
    pc = _chk(Pc, "Pc", torch.bfloat16) if Pc is not None else None
    _call("vu_softmax_stats", _chk(S, "S"), pc, B, h, N, ld, scale, drop_p, seed, sid, _chk(sums, "sums", torch.float64),
          int(precision), _stream(), nbytes=(4.0 + (2.0 if Pc is not None else 4.0)) * B * h * N * N)


def reattn_mix_reduce(P, dA, A, fold, B, h, N, ld, drop_p, seed, sid, red, tf32=False):
    pd, bf = _map(dA, "dA")
    pa, bf2 = _map(A, "A") if A is not None else (None, bf)       # A None: reductions only (tensor-core path)
    if bf != bf2:
        raise VuError("reattn_mix_reduce: A and dA must have the same dtype")
    pp, fmt, pb = _pmap(P, bf, tf32)
    eb = 2.0 if bf else 4.0
    _call("vu_reattn_mix_reduce", pp, pd, pa, fmt, _chk(fold, "fold"), B, h, N, ld,
          drop_p, seed, sid, _chk(red, "red", torch.float64), _stream(),
          nbytes=(pb + eb * (2 if A is not None else 1)) * B * h * N * N,
          flops=(4.0 if A is not None else 2.0) * h * B * h * N * N)


def reattn_stats(P, B, h, N, ld, drop_p, seed, sid, sums):
    _call("vu_reattn_stats", _chk(P, "P"), B, h, N, ld, drop_p, seed, sid, _chk(sums, "sums", torch.float64), _stream(),
          nbytes=4.0 * B * h * N * N)


def reattn_bn_finalize(sums, count, h, N, W, bconv, gamma, beta, rmean, rvar, nbt, eps, momentum, train,
                       fold, saved):
    call("vu_reattn_bn_finalize", _opt(sums, "sums", torch.float64), count, h, N, _chk(W, "W"),
         _chk(bconv, "bconv"), _chk(gamma, "gamma"), _chk(beta, "beta"), _chk(rmean, "running_mean"),
         _chk(rvar, "running_var"), _opt(nbt, "num_batches_tracked", torch.int64), eps, momentum, int(train),
         _chk(fold, "fold"), _chk(saved, "saved"), _stream())


def reattn_mix(P, A, fold, B, h, N, ld, drop_p, seed, sid, tf32=False):
    pa, bf = _map(A, "A")
    pp, fmt, pb = _pmap(P, bf, tf32)
    _call("vu_reattn_mix", pp, pa, fmt, _chk(fold, "fold"), B, h, N, ld, drop_p, seed, sid, _stream(),
          nbytes=(pb + (2.0 if bf else 4.0)) * B * h * N * N, flops=2.0 * h * B * h * N * N)


def reattn_bwd_reduce(P, dA, B, h, N, ld, drop_p, seed, sid, red):
    _call("vu_reattn_bwd_reduce", _chk(P, "P"), _chk(dA, "dA"), B, h, N, ld, drop_p, seed, sid,
          _chk(red, "red", torch.float64), _stream(), nbytes=2 * 4.0 * B * h * N * N)


def reattn_bwd_params(red, sums, B, h, N, W, bconv, gamma, saved, train, coef, dW, dbconv, dgamma, dbeta):
    call("vu_reattn_bwd_params", _chk(red, "red", torch.float64), _opt(sums, "sums", torch.float64), B, h, N,
         _chk(W, "W"), _chk(bconv, "bconv"), _chk(gamma, "gamma"), _chk(saved, "saved"), int(train),
         _chk(coef, "coef"), _chk(dW, "dW"), _chk(dbconv, "dbconv"), _chk(dgamma, "dgamma"), _chk(dbeta, "dbeta"),
         _stream())


def reattn_bwd_rows(P, dA, B, h, N, ld, W, bconv, gamma, saved, coef, train, scale, drop_p, seed, sid, tf32=False):
    pd, bf = _map(dA, "dA")
    pp, fmt, pb = _pmap(P, bf, tf32)
    _call("vu_reattn_bwd_rows", pp, pd, fmt, B, h, N, ld, _chk(W, "W"), _chk(bconv, "bconv"),
          _chk(gamma, "gamma"), _chk(saved, "saved"), _chk(coef, "coef"), int(train), scale, drop_p, seed, sid,
          _stream(), nbytes=(pb + 2 * (2.0 if bf else 4.0)) * B * h * N * N, flops=4.0 * h * B * h * N * N)


# ----------------------------------------------------------------------------------------- streamed re-attention
STREAM_EVAL, STREAM_STATS, STREAM_APPLY = 0, 1, 2


def stream_mask_bytes(B: int, N: int) -> int:
    """size of the cached dropout keep-bit buffer: 64 bits per lane of every (image, 16-row unit, 16-key step),
    whatever the head count (bit 4*head + key; 4-head models use the low half)"""
    return B * (N // 16) * (N // 16) * 32 * 8


def reattn_stream_supported(h: int, hd: int, N: int) -> bool:
    return bool(_lib.load().vu_reattn_stream_supported(h, hd, N))


def reattn_stream_fwd(mode, q, k, vt, o, fold, rowc, sums, pc, B, h, N, hd, scale, drop_p=0.0, seed=0, sid=0,
                      mask=None, amap=None):
    """Streamed Re-Attention forward (no (B,h,N,N) map): see include/vit_unet_b200.h vu_reattn_stream_fwd."""
    ldn = vt.shape[-1] if vt is not None else 0
    sweeps = {STREAM_EVAL: 2, STREAM_STATS: 2, STREAM_APPLY: 1}[mode]
    pv = 0 if mode == STREAM_STATS else 1
    _call("vu_reattn_stream_fwd", mode, _chk(q, "q"), _chk(k, "k"), _opt(vt, "vt", torch.bfloat16), _opt(o, "o"),
          _opt(fold, "fold"), _opt(rowc, "rowc"), _opt(sums, "sums", torch.float64), _opt(pc, "pc", torch.bfloat16),
          _opt(amap, "amap", torch.bfloat16), _opt(mask, "mask", torch.uint8), B, h, N, hd, ldn, scale, drop_p, seed, sid, _stream(),
          flops=2.0 * B * h * N * N * hd * (sweeps + pv) + (2.0 * h * B * h * N * N if pv else 0.0),
          nbytes=4.0 * B * N * h * hd * (2 + pv) + 2.0 * B * N * h * hd * pv + (2.0 * B * h * N * N if pc is not None else 0.0) + (2.0 * B * h * N * N if amap is not None else 0.0))


def reattn_stream_bwd_reduce(pc, mask, dO, v, red, B, h, N, hd, drop_p, seed, sid):
    _call("vu_reattn_stream_bwd_reduce", _chk(pc, "pc", torch.bfloat16), _opt(mask, "mask", torch.uint8), _chk(dO, "dO"),
          _chk(v, "v"), _chk(red, "red", torch.float64), B, h, N, hd, drop_p, seed, sid, _stream(),
          flops=2.0 * B * h * N * N * hd + 2.0 * h * B * h * N * N, nbytes=2.0 * B * h * N * N + 8.0 * B * N * h * hd)


def reattn_stream_bwd_ds(pc, mask, dO, v, kt, dS, dq, W, bconv, gamma, saved, coef, train, B, h, N, hd, drop_p, seed, sid):
    _call("vu_reattn_stream_bwd_ds", _chk(pc, "pc", torch.bfloat16), _opt(mask, "mask", torch.uint8), _chk(dO, "dO"),
          _chk(v, "v"), _chk(kt, "kt", torch.bfloat16), _chk(dS, "dS", torch.bfloat16), _chk(dq, "dq"), _chk(W, "W"),
          _chk(bconv, "bconv"), _chk(gamma, "gamma"), _chk(saved, "saved"), _opt(coef, "coef"), int(train), B, h, N, hd,
          kt.shape[-1], drop_p, seed, sid, _stream(),
          flops=4.0 * B * h * N * N * hd + 4.0 * h * B * h * N * N, nbytes=(2.0 * 2 + 2.0 * 3) * B * h * N * N + 14.0 * B * N * h * hd)


# ----------------------------------------------------------------------------------------- layer norm
LN_SCRATCH = 2 + 2 * _lib.LN_SPLIT      # floats of scratch per image for ln_stats / ln_bwd


def ln_stats(x, B, n, eps, stats, scratch=None):
    if scratch is None:
        scratch = torch.empty(B * LN_SCRATCH, dtype=torch.float32, device=x.device)
    _call("vu_ln_stats", _chk(x, "x"), B, n, eps, _chk(stats, "stats"), _chk(scratch, "scratch"), _stream(),
          nbytes=4.0 * B * n)


def ln_apply(x, stats, w, b, out, B, n, out16=None):
    """out16: optional bf16 copy of the result (bf16 mode: the next GEMM's A operand)"""
    _call("vu_ln_apply", _chk(x, "x"), _chk(stats, "stats"), _chk(w, "w"), _chk(b, "b"), _chk(out, "out"),
          _opt(out16, "out16", torch.bfloat16), B, n, _stream(), nbytes=(8.0 + (2.0 if out16 is not None else 0.0)) * B * n + 8.0 * n)


def ln_bwd(g, x, stats, w, dx, dw, db, scratch, B, n, dx16=None):
    _call("vu_ln_bwd", _chk(g, "g"), _chk(x, "x"), _chk(stats, "stats"), _chk(w, "w"), _chk(dx, "dx"),
          _opt(dx16, "dx16", torch.bfloat16), _chk(dw, "dw"), _chk(db, "db"), _chk(scratch, "scratch"), B, n, _stream(),
          nbytes=(12.0 + (2.0 if dx16 is not None else 0.0)) * B * n + 12.0 * n)


# ----------------------------------------------------------------------------------------- losses / misc
def loss_fwd(kind, pred, target, sums, loss):
    _call("vu_loss_fwd", LOSS_KINDS[kind], _chk(pred, "pred"), _chk(target, "target"), pred.numel(),
          _chk(sums, "sums", torch.float64), _chk(loss, "loss"), _stream(), nbytes=8.0 * pred.numel())


def loss_finalize(kind, n, sums, loss):
    _call("vu_loss_finalize", LOSS_KINDS[kind], n, _chk(sums, "sums", torch.float64), _chk(loss, "loss"), _stream())


def loss_bwd(kind, pred, target, sums, gscale, dpred):
    _call("vu_loss_bwd", LOSS_KINDS[kind], _chk(pred, "pred"), _chk(target, "target"), pred.numel(),
          _chk(sums, "sums", torch.float64), _chk(gscale, "gscale"), _chk(dpred, "dpred"), _stream(),
          nbytes=12.0 * pred.numel())


def psnr(pred, target, data_range=0.0):
    """Per-image PSNR (B,) on the device; data_range <= 0 selects skimage's float rule (1 if min(target) >= 0 else 2)."""
    B = pred.shape[0]
    n = pred.numel() // B
    scratch = torch.empty(2 * B, dtype=torch.float64, device=pred.device)
    out = torch.empty(B, dtype=torch.float32, device=pred.device)
    _call("vu_psnr", _chk(pred, "pred"), _chk(target, "target"), B, n, float(data_range),
          _chk(scratch, "scratch", torch.float64), _chk(out, "psnr"), _stream(), nbytes=12.0 * pred.numel())
    return out


def u8hwc_to_chw(src, scale=1.0 / 255.0, mean=0.0, std=1.0):
    """uint8 (B,H,W,C) -> float32 (B,C,H,W), (src*scale - mean)/std."""
    B, H, W, Cc = src.shape
    dst = torch.empty((B, Cc, H, W), dtype=torch.float32, device=src.device)
    _call("vu_u8hwc_to_chw", _chk(src, "src", torch.uint8), _chk(dst, "dst"), B, Cc, H, W, scale, mean, std, _stream(),
          nbytes=5.0 * src.numel())
    return dst


def resize_u8hwc(src, Hd, Wd):
    """cv2.resize(INTER_LINEAR) of a uint8 (B,H,W,C) batch on the device."""
    B, Hs, Ws, Cc = src.shape
    dst = torch.empty((B, Hd, Wd, Cc), dtype=torch.uint8, device=src.device)
    _call("vu_resize_u8hwc", _chk(src, "src", torch.uint8), _chk(dst, "dst", torch.uint8), B, Cc, Hs, Ws, Hd, Wd, _stream(),
          nbytes=float(src.numel() + dst.numel()))
    return dst


def warp_u8hwc_to_chw(src, mats, Hd, Wd, bilinear=True, border=0.0, round_u8=True, scale=1.0 / 255.0, mean=0.0, std=1.0,
                      post=1.0):
    """uint8 (B,H,W,C) -> float32 (B,C,Hd,Wd) through per-image inverse affine maps `mats` (B,6) or the identity."""
    B, Hs, Ws, Cc = src.shape
    dst = torch.empty((B, Cc, Hd, Wd), dtype=torch.float32, device=src.device)
    _call("vu_warp_u8hwc_to_chw", _chk(src, "src", torch.uint8), _chk(dst, "dst"), _opt(mats, "mats"), B, Cc, Hs, Ws, Hd, Wd,
          int(bilinear), border, int(round_u8), scale, mean, std, post, _stream(), nbytes=float(src.numel() + 4 * dst.numel()))
    return dst


def dropout(x, out, p, seed, sid):
    """out fp32, or bf16 (bf16 mode: masked gradient as a GEMM operand; p == 0 is a plain conversion)"""
    bf = out.dtype == torch.bfloat16
    _call("vu_dropout", _chk(x, "in"), _chk(out, "out", out.dtype if bf else torch.float32), int(bf), x.numel(), p, seed, sid,
          _stream(), nbytes=(6.0 if bf else 8.0) * x.numel())
    return out


def cast_bf16(w, want_t=True, want_n=True):
    """fp32 (R,C) -> (bf16 (R,C) or None, bf16 (C,R) or None): per-step copies of a Linear weight for the bf16 mode"""
    R, Cc = w.shape
    dn = torch.empty((R, Cc), dtype=torch.bfloat16, device=w.device) if want_n else None
    dt = torch.empty((Cc, R), dtype=torch.bfloat16, device=w.device) if want_t else None
    _call("vu_cast_bf16", _chk(w, "w"), _opt(dn, "dst", torch.bfloat16), _opt(dt, "dst_t", torch.bfloat16), R, Cc, _stream(),
          nbytes=(4.0 + 2.0 * (int(want_n) + int(want_t))) * R * Cc)
    return dn, dt


def axpby(x, y, a, b):
    _call("vu_axpby", _chk(x, "x"), _chk(y, "y"), x.numel(), a, b, _stream(), nbytes=12.0 * x.numel())
    return y


def adamw(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    call("vu_adamw", _chk(p, "p"), _chk(g, "g"), _chk(m, "m"), _chk(v, "v"), p.numel(), lr, beta1, beta2, eps,
         wd, step, grad_scale, _stream())


def gemm_tf32_fallbacks() -> int:
    """TF32 GEMM requests that ran on the CUDA-core kernel (operand not TMA-addressable); 0 on the benchmarked path."""
    return int(_lib.load().vu_gemm_tf32_fallbacks())


def sm_count(device: int = 0) -> int:
    return _lib.load().vu_device_sm_count(device)
