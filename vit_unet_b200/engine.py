"""Forward / backward schedule of the ViT-UNet hot path on the CUDA kernels.

This is the host side of the path: a fixed sequence of C-ABI launches per block (no tracing compiler, no
autograd graph inside the network -- the whole model is ONE autograd node, see model.py).  Reference order of
operations: HViT_UNet.forward (model.py:372-435) -> ReAttentionTransformerEncoder.forward (:201-207) ->
ReAttention.forward (:150-164) / SkipConnection.forward (:244-259) / FeedForward (:95-110).

Activations are fp32 token tensors (B, N_l, D_l) in the patch layout of their level; N_l * D_l = C*H*W at every
level, so "downsampling"/"upsampling" are permutations (vu_repatch).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops

_PRECISION = {"value": ops.PREC_FP32}
_CHUNK_STREAM = 4096          # dropout RNG stream offset per image slice (layer stream ids stay far below this)
# bytes of attention maps a slice may keep live between consecutive launches (two maps: S/P and A, or P-slice and
# dA/dS) when the batch is processed in image slices meant to stay L2-resident.  0 disables slicing (default):
# measured on B200 (Base, 64 images) slicing LOSES -- 80 MB: 822 img/s, 160 MB: 985 img/s, off: 1283 img/s -- the
# per-slice launches are too small to fill 148 SMs and the launch count grows 4x.  Kept for memory-bound inference.
_MAP_L2_BYTES = {"value": int(float(os.environ.get("VU_MAP_L2_MB", "0")) * (1 << 20))}


# bf16 storage of the mixed map A and the gradient map dA/dS (probabilities P stay fp32) on the tensor-core path:
# halves the HBM bytes of the five map-consuming GEMMs and of the map kernels' A / dA traffic.
# q/k/v convs as implicit GEMMs on TF32 warp MMAs (vu_conv.cu, tf32 / bf16 modes): OPT-IN (VU_CONV_MMA=1).  Measured on B200
# (Base step, 256 images, ms per step, MMA vs FFMA kernels): forward 3.20 vs 2.15, data gradient 9.92 vs 3.10, weight gradient
# 3.01 vs 3.03 -- the gathered fragments cost one LDS.32 per operand word and the accumulator layout scatters the stores into
# 32-byte sectors, so the LSU, not the FMA pipe, is the limit; and the extra TF32 rounding of x / dy costs parity margin on the
# tiny configs (7.8e-2 vs the 5e-2 asserted).  The FFMA patch kernels stay the default in every mode.
_CONV_MMA = {"value": os.environ.get("VU_CONV_MMA", "0") == "1"}
_BF16_MAPS = {"value": os.environ.get("VU_BF16_MAPS", "1") == "1"}      # on by default in the tf32 mode (1e-2 class)


# centred bf16 probabilities (train, tf32 mode): halves the saved-map memory; off by default (about +1% images/s only,
# the map kernels are issue-bound rather than HBM-bound once the mixing runs on the tensor cores)
_BF16_PROBS = {"value": os.environ.get("VU_BF16_PROBS", "0") == "1"}
_KEEP_MIXED_MAP = {"value": os.environ.get("VU_KEEP_MIXED_MAP", "1") == "1"}
_BF16_PROBS_LONG = {"value": os.environ.get("VU_BF16_PROBS_LONG", "1") == "1"}     # centred bf16 probabilities where N >= 256


# streamed Re-Attention (vu_reattn_stream.cu) on the tensor-core path where the shape is supported: no (B,h,N,N) map is
# written in inference; in training only the centred bf16 probabilities, the mixed map and dS cross HBM.  OPT-IN
# (set_streamed / VU_STREAMED=1): measured on B200 (tools/block_bench.py, 256 images, Base level-2 block) the streamed
# kernels run at 7 warps per SM (all heads of a position live in one lane: 240-255 registers) and are latency-bound --
# forward 4.9 ms vs 4.4 ms materialised, backward 12.6 vs 5.7 ms -- while using 40 % less memory (7.6 vs 12.3 GB).
_STREAMED = {"value": os.environ.get("VU_STREAMED", "0") == "1"}
_STREAMED_BWD = {"value": os.environ.get("VU_STREAMED_BWD", "1") == "1"}     # with it: the streamed backward kernels too
# INFERENCE (nothing saved for backward) is different: there the streamed kernel replaces fp32 map round trips of up to
# 157 MB per image and block (Lite level 2: 3136 tokens) and needs no batch slicing, so it is on by default
_STREAMED_INFER = {"value": os.environ.get("VU_STREAMED_INFER", "1") == "1"}


def set_streamed(on: bool, backward: bool = True, inference=None) -> None:
    """Streamed Re-Attention for training steps (forward, and with `backward` the backward kernels); `inference`
    (default: unchanged) switches the streamed no-grad forward, which is on by default."""
    _STREAMED["value"] = bool(on)
    _STREAMED_BWD["value"] = bool(on and backward)
    if inference is not None:
        _STREAMED_INFER["value"] = bool(inference)


def set_bf16_maps(on: bool) -> None:
    _BF16_MAPS["value"] = bool(on)


def set_bf16_probs(on: bool, long_rows=None) -> None:
    """Store the train-mode attention probabilities as centred bf16 (P - 1/N) wherever the tensor-core map path applies
    (`on`), or only where rows are long (`long_rows`, N >= 256: the default)."""
    _BF16_PROBS["value"] = bool(on)
    if long_rows is not None:
        _BF16_PROBS_LONG["value"] = bool(long_rows)


def set_map_l2_budget(megabytes: float) -> None:
    _MAP_L2_BYTES["value"] = int(megabytes * (1 << 20))


PREC_BF16 = 2       # engine-level mode; at the C ABI the bf16-operand GEMMs are VU_PREC_TF32 requests with a_bf16 / b_bf16 set


def set_precision(mode: str) -> None:
    """'fp32': CUDA-core FMA everywhere (1e-5 parity).  'tf32': tcgen05 tensor cores for the contractions, fp32 storage.
    'bf16': the tensor-core path with bf16 storage and bf16 (kind::f16) tcgen05 products for the token GEMMs: every GEMM
    operand that is not the residual stream (attention output, LayerNorm output copy, FeedForward hidden, the gradients
    entering the data- and weight-gradient products) and the per-step copies of the Linear weights are bfloat16; the
    residual stream, LayerNorm / BatchNorm statistics, master weights and all gradients of parameters stay fp32."""
    _PRECISION["value"] = {"fp32": ops.PREC_FP32, "tf32": ops.PREC_TF32, "bf16": PREC_BF16}[mode]


def get_precision() -> str:
    return {ops.PREC_FP32: "fp32", ops.PREC_TF32: "tf32", PREC_BF16: "bf16"}[_PRECISION["value"]]


@dataclass
class Geometry:
    C: int
    S: int                 # square image side
    p0: int
    depth: int
    depth_te: int
    n_bottleneck: int
    heads: int
    hidden: int
    shared_ln: bool        # README variant: one LN per block used twice (ViT_UNet.ipynb c27)
    pe_conv: bool          # README variant applies conv2d in the PatchEncoder (ViT_UNet.ipynb c16:L32-33)
    table_p: int           # patch size the position table is indexed at (p0: model.py:80-82; p0/2^depth: c16)
    out_conv: bool
    attn_drop: float = 0.0
    proj_drop: float = 0.0
    linear_drop: float = 0.0

    def p(self, l): return self.p0 // (2 ** l)
    def N(self, l): return (self.S // self.p(l)) ** 2
    def D(self, l): return self.C * self.p(l) ** 2
    def Hd(self, l): return self.hidden // (2 ** l)

    def schedule(self):
        """Forward order of blocks: ('block', prefix, level) / ('down', l) / ('up', l) / ('skip', prefix, level)."""
        s = []
        i = 0
        for l in range(self.depth):
            for _ in range(self.depth_te):
                s.append(("block", f"Encoders.{i}.", l)); i += 1
            s.append(("down", l))
        for i in range(self.n_bottleneck):
            s.append(("block", f"BottleNeck.{i}.", self.depth))
        i = 0
        for lv in range(self.depth):
            l = self.depth - lv
            for _ in range(self.depth_te):
                s.append(("block", f"Decoders.{i}.", l)); i += 1
            s.append(("up", l))
            s.append(("skip", f"SkipConnections.{lv}.", l - 1))
        return s

    def param_order(self, names: List[str]) -> List[str]:
        """Parameter names in forward-execution order (backward finishes them in exactly the reverse order, so
        a flat gradient buffer in this order becomes ready as a growing suffix -> contiguous all-reduce buckets)."""
        groups = ["PE."]
        for st in self.schedule():
            if st[0] in ("block", "skip"):
                groups.append(st[1])
        groups.append("conv2d.")
        out = []
        for gname in groups:
            out += [n for n in names if n.startswith(gname) and n not in out]
        assert sorted(out) == sorted(names), "parameter naming drifted from the schedule"
        return out


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


class Engine:
    """Runs the schedule.  `P` maps state_dict-style names to CUDA tensors (parameters and BN buffers)."""

    def __init__(self, geom: Geometry):
        self.g = geom
        self.sched = geom.schedule()
        self.on_grads_ready = None     # callback(first_param_index) used by the data-parallel wrapper
        self.precision = None          # per-module override of the global precision mode (ViT_UNet(dtype=torch.bfloat16))
        self._w16 = {}                 # bf16 copies of the Linear weights of the forward pass in flight (bf16 mode)

    # ------------------------------------------------------------------------------------------- helpers
    def _mode(self) -> int:
        return self.precision if self.precision is not None else _PRECISION["value"]

    def _prec(self) -> int:
        """precision class handed to the C ABI: the bf16 mode runs on the tensor-core (VU_PREC_TF32) entry points"""
        m = self._mode()
        return ops.PREC_TF32 if m == PREC_BF16 else m

    def _b16(self) -> bool:
        return self._mode() == PREC_BF16

    def _gemm_tokens(self, A, W, out, M, N, K, **kw):
        """out[M,N] = A[M,K] @ W[N,K]^T (+epilogue): the nn.Linear shape."""
        return ops.gemm(A, W, out, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, precision=self._prec(), **kw)

    def _weight16(self, P, name, transposed=False):
        """bf16 copy of Linear weight `name` (out,in), or of its transpose (in,out): made once per forward pass, kept with
        the saved tensors for the backward pass (the master weight stays fp32)"""
        ent = self._w16.get(name)
        if ent is None:
            ent = ops.cast_bf16(P[name], want_t=self._w16_need_t, want_n=True)
            self._w16[name] = ent
        if transposed and ent[1] is None:
            ent = (ent[0], ops.cast_bf16(P[name], want_t=True, want_n=False)[1])
            self._w16[name] = ent
        return ent[1] if transposed else ent[0]

    @staticmethod
    def _map_chunk(B, h, N, ld):
        """Images per slice so that two (c,h,N,ld) fp32 maps fit the L2 budget."""
        budget = _MAP_L2_BYTES["value"]
        if budget <= 0:
            return B
        return max(1, min(B, budget // (2 * h * N * ld * 4)))

    # ------------------------------------------------------------------------------------------- attention
    def _attn_fwd(self, P, pre, xq, xkv, l, B, train, seed, sid, residual, saved):
        g = self.g
        N, D, h, p = g.N(l), g.D(l), g.heads, g.p(l)
        hd, ld = D // h, ops.pad4(N)
        prec, b16 = self._prec(), self._b16()
        wq, wk, wv = P[pre + "qconv2d.weight"], P[pre + "kconv2d.weight"], P[pre + "vconv2d.weight"]
        q, k, v = _empty((B, N, D), xq), _empty((B, N, D), xq), _empty((B, N, D), xq)
        tc = prec == ops.PREC_TF32 and _CONV_MMA["value"]     # tensor-core class: q/k/v convs as implicit GEMMs on warp MMAs
        if xq is xkv:         # the three filters are uploaded from their own parameter tensors (no concatenation pass)
            ops.conv3x3_fwd(xq, p, [wq, wk, wv], None, [q, k, v], p, p, B, g.C, g.S, g.S, tf32=tc)
        else:
            ops.conv3x3_fwd(xq, p, wq, None, [q], p, p, B, g.C, g.S, g.S, tf32=tc)
            ops.conv3x3_fwd(xkv, p, [wk, wv], None, [k, v], p, p, B, g.C, g.S, g.S, tf32=tc)
        # The (B,h,N,N) maps are processed in slices of `c` images sized so that the maps a kernel chain hands from
        # one launch to the next (S -> P -> A, ~2 live maps) stay resident in the 126 MB L2: HBM then sees P once on
        # the way out (it is saved for backward) and once on the way back in, instead of ~6 full passes.
        scale = float(hd) ** -0.5
        Wm = P[pre + "reatten_matrix.weight"].reshape(h, h)
        bm = P[pre + "reatten_matrix.bias"]
        adrop = g.attn_drop if train else 0.0
        c = self._map_chunk(B, h, N, ld)
        keep_P = saved is not None
        bf16 = prec == ops.PREC_TF32 and _BF16_MAPS["value"] and N % 8 == 0
        # streamed forward: always for inference; with a backward pass to feed only where the materialised backward
        # kernels accept centred bf16 probabilities (8 heads, bf16 maps)
        # (4-head models in EVAL mode with a backward pass to feed stay on the exact fp32 map kernels: with running-statistics
        # BatchNorm over 3136-key rows the score gradient is a 1e-3 residue of cancelling terms, below what bf16 map storage
        # resolves -- measured O(1) relative errors on Lite's eval-mode gradients; train mode is unaffected)
        stream_on = ((_STREAMED["value"] and (train or not keep_P or h == 8))
                     or (_STREAMED_INFER["value"] and not keep_P and not train))
        # with a backward pass to feed: either the streamed backward kernels take over (any supported head count), or the
        # materialised tensor-core backward reads the centred bf16 probabilities (8 heads only)
        bwd_ok = (not keep_P or _STREAMED_BWD["value"] or (bf16 and ops.reattn_tensor_core_path(h, N, ld)))
        if prec == ops.PREC_TF32 and stream_on and c >= B and ops.reattn_stream_supported(h, hd, N) and bwd_ok:
            return self._attn_fwd_streamed(P, pre, xq, xkv, q, k, v, l, B, train, seed, sid, residual, saved)
        # train mode on the tensor-core map path: probabilities are kept as CENTRED bf16 (P - 1/N), scores are a
        # per-slice fp32 scratch -- the saved map and every later pass over it cost half the bytes
        # centred bf16 probabilities: on request everywhere the tensor-core map kernels run, and by default for long rows
        # (N >= 256: the BatchNorm statistics average over >= 65k values per image and head; measured +0.7 .. 1.4 % images/s
        # and half the saved-map memory; the few-token levels keep fp32 probabilities).
        # Negative result kept out of the tree: caching the dropout keep-bits (one byte per key quad, written by the
        # statistics kernel, read by mix / mix_reduce / bwd_rows instead of re-hashing) LOST 3.6 % -- the scattered byte
        # loads cost these HBM-bound kernels more than the ~25 integer instructions of the hash (93.3 vs 90.0 ms per step)
        pc16 = (bf16 and train and ops.reattn_tensor_core_path(h, N, ld)
                and (_BF16_PROBS["value"] or (_BF16_PROBS_LONG["value"] and N >= 256)))
        if pc16:
            Pm = torch.empty((B, h, N, ld), dtype=torch.bfloat16, device=xq.device)
            Sc = _empty((c, h, N, ld), xq)
        else:
            Pm = _empty((B if (keep_P or train) else c, h, N, ld), xq)
        A = torch.empty((c, h, N, ld), dtype=torch.bfloat16 if bf16 else torch.float32, device=xq.device)
        # bf16 mode: the attention output is only ever a GEMM operand (proj forward, proj weight gradient) -> bf16
        O = torch.empty((B, N, D), dtype=torch.bfloat16, device=xq.device) if b16 else _empty((B, N, D), xq)
        vt = ops.heads_transpose_bf16(v, B, N, D, h) if bf16 else None        # (B,h,hd,ldn): K-major B operand of A.V
        ldn = vt.shape[-1] if bf16 else 0
        fold, bn_saved = _empty((h * h + h,), xq), _empty((2 * h,), xq)
        sums = ops.zeros(h + h * h, torch.float64, xq.device) if train else None

        def scores(b0, bc, dst):
            ops.gemm(q[b0:b0 + bc], k[b0:b0 + bc], dst, N, N, hd, trans_b=True, lda=D, ldb=D, ldc=ld, batch_outer=bc,
                     batch_inner=h, sA=(N * D, hd), sB=(N * D, hd), sC=(h * N * ld, N * ld), precision=prec)

        def mix_pv(b0, bc, src, ci):
            ops.reattn_mix(src, A[:bc], fold, bc, h, N, ld, adrop, seed, sid + _CHUNK_STREAM * ci,
                           tf32=prec == ops.PREC_TF32)
            if bf16:
                ops.gemm(A[:bc], vt[b0:b0 + bc], O[b0:b0 + bc], N, hd, N, trans_b=True, lda=ld, ldb=ldn, ldc=D,
                         batch_outer=bc, batch_inner=h, sA=(h * N * ld, N * ld), sB=(h * hd * ldn, hd * ldn),
                         sC=(N * D, hd), precision=prec)
            else:
                ops.gemm(A[:bc], v[b0:b0 + bc], O[b0:b0 + bc], N, hd, N, trans_b=False, lda=ld, ldb=D, ldc=D,
                         batch_outer=bc, batch_inner=h, sA=(h * N * ld, N * ld), sB=(N * D, hd), sC=(N * D, hd),
                         precision=prec)

        def finalize():
            ops.reattn_bn_finalize(sums, B * N * N, h, N, Wm, bm, P[pre + "var_norm.weight"], P[pre + "var_norm.bias"],
                                   P[pre + "var_norm.running_mean"], P[pre + "var_norm.running_var"],
                                   P.get(pre + "var_norm.num_batches_tracked"), 1e-5, 0.1, train, fold, bn_saved)

        if train:
            # batch statistics couple all images: phase 1 (scores, softmax, moments) per slice, then phase 2
            for ci, b0 in enumerate(range(0, B, c)):
                bc = min(c, B - b0)
                if pc16:
                    scores(b0, bc, Sc[:bc])
                    ops.softmax_stats(Sc[:bc], bc, h, N, ld, scale, adrop, seed, sid + _CHUNK_STREAM * ci, sums,
                                      precision=prec, Pc=Pm[b0:b0 + bc])
                else:
                    scores(b0, bc, Pm[b0:b0 + bc])
                    ops.softmax_stats(Pm[b0:b0 + bc], bc, h, N, ld, scale, adrop, seed, sid + _CHUNK_STREAM * ci, sums,
                                      precision=prec)
            finalize()
            for ci, b0 in enumerate(range(0, B, c)):
                bc = min(c, B - b0)
                mix_pv(b0, bc, Pm[b0:b0 + bc], ci)
        else:
            finalize()                                   # running statistics: the whole chain runs per slice
            for ci, b0 in enumerate(range(0, B, c)):
                bc = min(c, B - b0)
                dst = Pm[b0:b0 + bc] if keep_P else Pm[:bc]
                scores(b0, bc, dst)
                ops.softmax_rows(dst, bc * h * N, N, ld, scale)
                mix_pv(b0, bc, dst, ci)
        # one slice on the tensor-core map path: keep the mixed map for backward (dV = A^T dO) instead of recomputing
        # and re-writing it there -- 2 bytes per map element of extra memory for one pass less over the maps
        keep_A = (saved is not None and c >= B and _KEEP_MIXED_MAP["value"] and (bf16 or prec == ops.PREC_TF32)
                  and ops.reattn_tensor_core_path(h, N, ld))
        if not keep_A:
            A = None
        if not keep_P:
            del Pm
            Pm = None
        y = _empty((B, N, D), xq)
        pdrop = g.proj_drop if train else 0.0
        Wp = self._weight16(P, pre + "proj.weight") if b16 else P[pre + "proj.weight"]
        self._gemm_tokens(O, Wp, y, B * N, D, D, bias=P[pre + "proj.bias"],
                          residual=residual, drop_p=pdrop, drop_seed=seed, drop_stream=sid + 1)
        if saved is not None:
            saved.update(xq=xq, xkv=xkv, q=q, k=k, v=v, Pm=Pm, O=O, fold=fold, bn=bn_saved, sums=sums, seed=seed, sid=sid,
                         adrop=adrop, pdrop=pdrop, train=train, chunk=c, bf16=bf16, A=A)
        return y

    def _attn_fwd_streamed(self, P, pre, xq, xkv, q, k, v, l, B, train, seed, sid, residual, saved):
        """softmax -> dropout -> head mixing + BatchNorm -> A.V without the (B,h,N,N) maps (vu_reattn_stream.cu)."""
        g = self.g
        N, D, h = g.N(l), g.D(l), g.heads
        hd = D // h
        scale = float(hd) ** -0.5
        adrop = g.attn_drop if train else 0.0
        keep_P = saved is not None
        vt = ops.heads_transpose_bf16(v, B, N, D, h)
        fold, bn_saved = _empty((h * h + h,), xq), _empty((2 * h,), xq)
        sums = ops.zeros(h + h * h, torch.float64, xq.device) if (train or keep_P) else None
        O = _empty((B, N, D), xq)
        Pm = torch.empty((B, h, N, N), dtype=torch.bfloat16, device=xq.device) if keep_P else None
        A, mask = None, None

        def finalize():
            ops.reattn_bn_finalize(sums if train else None, B * N * N, h, N, P[pre + "reatten_matrix.weight"].reshape(h, h),
                                   P[pre + "reatten_matrix.bias"], P[pre + "var_norm.weight"], P[pre + "var_norm.bias"],
                                   P[pre + "var_norm.running_mean"], P[pre + "var_norm.running_var"],
                                   P.get(pre + "var_norm.num_batches_tracked"), 1e-5, 0.1, train, fold, bn_saved)
        if train or keep_P:
            rowc = _empty((B, h, N), xq)
            # keep-bits of the dropout mask: generated (hashed) once by the statistics launch, re-read by the apply launch
            mask = torch.empty(ops.stream_mask_bytes(B, N), dtype=torch.uint8, device=xq.device) if adrop > 0 else None
            ops.reattn_stream_fwd(ops.STREAM_STATS, q, k, None, None, None, rowc, sums, Pm, B, h, N, hd, scale, adrop, seed, sid,
                                  mask=mask)
            finalize()
            # the mixed map is kept (bf16) for the backward product dV = A^T dO: one write here instead of a recompute there
            A = torch.empty((B, h, N, N), dtype=torch.bfloat16, device=xq.device) if keep_P else None
            ops.reattn_stream_fwd(ops.STREAM_APPLY, q, k, vt, O, fold, rowc, None, None, B, h, N, hd, scale, adrop, seed, sid,
                                  mask=mask, amap=A)
        else:
            finalize()
            ops.reattn_stream_fwd(ops.STREAM_EVAL, q, k, vt, O, fold, None, None, None, B, h, N, hd, scale)
        y = _empty((B, N, D), xq)
        pdrop = g.proj_drop if train else 0.0
        Wp = P[pre + "proj.weight"]
        if self._b16():       # the streamed kernels write fp32 rows: one conversion pass (6 bytes per token element)
            O = ops.dropout(O, torch.empty((B, N, D), dtype=torch.bfloat16, device=xq.device), 0.0, 0, 0)
            Wp = self._weight16(P, pre + "proj.weight")
        self._gemm_tokens(O, Wp, y, B * N, D, D, bias=P[pre + "proj.bias"],
                          residual=residual, drop_p=pdrop, drop_seed=seed, drop_stream=sid + 1)
        if saved is not None:
            saved.update(xq=xq, xkv=xkv, q=q, k=k, v=v, Pm=Pm, O=O, fold=fold, bn=bn_saved, sums=sums if train else None,
                         seed=seed, sid=sid, adrop=adrop, pdrop=pdrop, train=train, chunk=B, bf16=True, A=A, streamed=True,
                         mask=mask)
        return y

    def _attn_bwd(self, P, G, pre, dy, l, B, sv, dxq_acc, dxkv_acc, acc_q=True, acc_kv=True):
        """dy: grad of proj output (pre-residual).  Accumulates into dxq_acc / dxkv_acc (may be the same tensor); with
        acc_* False the conv data gradient is WRITTEN instead (skip connection: no zero-filled accumulator needed)."""
        g = self.g
        N, D, h, p = g.N(l), g.D(l), g.heads, g.p(l)
        hd, ld = D // h, ops.pad4(N)
        prec, b16 = self._prec(), self._b16()
        M = B * N
        q, k, v, Pm, O = sv["q"], sv["k"], sv["v"], sv["Pm"], sv["O"]
        seed, sid, adrop, pdrop, train = sv["seed"], sv["sid"], sv["adrop"], sv["pdrop"], sv["train"]
        if b16:            # masked gradient straight to bf16: the A operand of dO = dyd Wp and of dWp = dyd^T O
            dyd = ops.dropout(dy, torch.empty(dy.shape, dtype=torch.bfloat16, device=dy.device), pdrop, seed, sid + 1)
        elif pdrop > 0:
            dyd = ops.dropout(dy, torch.empty_like(dy), pdrop, seed, sid + 1)
        else:
            dyd = dy
        # proj: dO = dyd @ Wp ; dWp = dyd^T @ O ; dbp = colsum(dyd)
        dO = _empty((B, N, D), dy)
        if b16:
            ops.gemm(dyd, self._weight16(P, pre + "proj.weight", transposed=True), dO, M, D, D, trans_b=True, lda=D, ldb=D,
                     ldc=D, precision=prec)
        else:
            ops.gemm(dyd, P[pre + "proj.weight"], dO, M, D, D, trans_b=False, lda=D, ldb=D, ldc=D, precision=prec)
        self._wgrad(dyd, O, G[pre + "proj.weight"], M, D, D)
        ops.colsum(dyd, M, D, D, G[pre + "proj.bias"], accumulate=True)
        del dyd
        self._notify(pre + "proj.")        # [proj.weight, end of the flat gradient buffer) is final: an early all-reduce bucket
        Wm = P[pre + "reatten_matrix.weight"].reshape(h, h)
        bm = P[pre + "reatten_matrix.bias"]
        if sv.get("streamed") and _STREAMED_BWD["value"]:
            # streamed backward (vu_reattn_stream.cu): dA = dO v^T is formed on the fly in both kernels and never stored;
            # the maps that cross HBM are the saved centred probabilities (read), the saved mixed map (dV = A^T dO) and dS
            # (written once, read by dK = dS^T q)
            gamma = P[pre + "var_norm.weight"]
            mask, A_kept = sv.get("mask"), sv["A"]
            red = ops.zeros(h + h * h, torch.float64, dy.device)
            ops.reattn_stream_bwd_reduce(Pm, mask, dO, v, red, B, h, N, hd, adrop, seed, sid)
            dq, dk, dv = _empty((B, N, D), dy), _empty((B, N, D), dy), _empty((B, N, D), dy)
            dOt = ops.heads_transpose_bf16(dO, B, N, D, h)
            ldn = dOt.shape[-1]

            def map_gemm_t(Amap, Bt, Cout):          # Cout (B,N,D head-sliced) = Amap^T @ tokens of the head
                ops.gemm(Amap, Bt, Cout, N, hd, N, trans_a=True, trans_b=True, lda=N, ldb=ldn, ldc=D, batch_outer=B,
                         batch_inner=h, sA=(h * N * N, N * N), sB=(h * hd * ldn, hd * ldn), sC=(N * D, hd), precision=prec)
            map_gemm_t(A_kept, dOt, dv)
            sv["A"] = None
            del A_kept, dOt
            coef = _empty((2 * h,), dy)
            ops.reattn_bwd_params(red, sv["sums"], B, h, N, Wm, bm, gamma, sv["bn"], train, coef,
                                  G[pre + "reatten_matrix.weight"], G[pre + "reatten_matrix.bias"],
                                  G[pre + "var_norm.weight"], G[pre + "var_norm.bias"])
            kt = ops.heads_transpose_bf16(k, B, N, D, h)
            dS = torch.empty((B, h, N, N), dtype=torch.bfloat16, device=dy.device)
            ops.reattn_stream_bwd_ds(Pm, mask, dO, v, kt, dS, dq, Wm, bm, gamma, sv["bn"], coef, train, B, h, N, hd,
                                     adrop, seed, sid)
            del kt, dO
            qt = ops.heads_transpose_bf16(q, B, N, D, h)
            map_gemm_t(dS, qt, dk)
            del dS, qt
            self._qkv_conv_bwd(P, G, pre, sv, dq, dk, dv, p, B, dxq_acc, dxkv_acc, acc_q, acc_kv)
            return
        # recompute the mixed map, then dV = A^T dO ; dA = dO V^T
        # Same image slices as forward (identical dropout RNG streams).  Phase A per slice: dA = dO V^T, one pass over
        # (P, dA) -> recomputed mixed map A + backward reductions, dV = A^T dO.  Then the closed-form parameter
        # gradients / BatchNorm-backward means.  Phase B per slice: dA again (a K = head_dim GEMM, cheaper than an HBM
        # round trip), dA -> dS in place, dQ = dS K, dK = dS^T Q.  Only P crosses HBM (twice).
        c = sv["chunk"]
        scale = float(hd) ** -0.5
        gamma = P[pre + "var_norm.weight"]
        bf16 = sv["bf16"]
        mdt = torch.bfloat16 if bf16 else torch.float32
        dA = ops.zeros((c, h, N, ld), mdt, dy.device) if ld != N else torch.empty((c, h, N, ld), dtype=mdt, device=dy.device)
        A_kept = sv.get("A")
        A = A_kept if A_kept is not None else torch.empty((c, h, N, ld), dtype=mdt, device=dy.device)
        if bf16:     # per-head transposed bf16 copies: the K-major B operands of dV = A^T dO, dQ = dS K, dK = dS^T Q
            dOt, kt, qt = (ops.heads_transpose_bf16(t, B, N, D, h) for t in (dO, k, q))
            ldn = dOt.shape[-1]

        def map_gemm(Amap, trans_a, Bt, Bf, Cout, b0, bc):
            """Cout[b0:b0+bc] (B,N,D head-sliced) = op(Amap[:bc]) @ (tokens of head): bf16 maps use the transposed copy Bt"""
            if bf16:
                ops.gemm(Amap[:bc], Bt[b0:b0 + bc], Cout[b0:b0 + bc], N, hd, N, trans_a=trans_a, trans_b=True, lda=ld,
                         ldb=ldn, ldc=D, batch_outer=bc, batch_inner=h, sA=(h * N * ld, N * ld),
                         sB=(h * hd * ldn, hd * ldn), sC=(N * D, hd), precision=prec)
            else:
                ops.gemm(Amap[:bc], Bf[b0:b0 + bc], Cout[b0:b0 + bc], N, hd, N, trans_a=trans_a, trans_b=False, lda=ld,
                         ldb=D, ldc=D, batch_outer=bc, batch_inner=h, sA=(h * N * ld, N * ld), sB=(N * D, hd),
                         sC=(N * D, hd), precision=prec)
        dq, dk, dv = _empty((B, N, D), dy), _empty((B, N, D), dy), _empty((B, N, D), dy)
        red = ops.zeros(h + h * h, torch.float64, dy.device)

        def grad_map(b0, bc):
            ops.gemm(dO[b0:b0 + bc], v[b0:b0 + bc], dA[:bc], N, N, hd, trans_b=True, lda=D, ldb=D, ldc=ld,
                     batch_outer=bc, batch_inner=h, sA=(N * D, hd), sB=(N * D, hd), sC=(h * N * ld, N * ld),
                     precision=prec)

        for ci, b0 in enumerate(range(0, B, c)):
            bc = min(c, B - b0)
            grad_map(b0, bc)
            ops.reattn_mix_reduce(Pm[b0:b0 + bc], dA[:bc], None if A_kept is not None else A[:bc], sv["fold"], bc, h, N, ld,
                                  adrop, seed, sid + _CHUNK_STREAM * ci, red, tf32=prec == ops.PREC_TF32)
            map_gemm(A, True, dOt if bf16 else None, dO, dv, b0, bc)
        sv["A"] = None
        del A, A_kept
        coef = _empty((2 * h,), dy)
        ops.reattn_bwd_params(red, sv["sums"], B, h, N, Wm, bm, gamma, sv["bn"], train, coef,
                              G[pre + "reatten_matrix.weight"], G[pre + "reatten_matrix.bias"],
                              G[pre + "var_norm.weight"], G[pre + "var_norm.bias"])
        single = c >= B           # one slice: dA from phase A is still intact
        for ci, b0 in enumerate(range(0, B, c)):
            bc = min(c, B - b0)
            if not single:
                grad_map(b0, bc)
            ops.reattn_bwd_rows(Pm[b0:b0 + bc], dA[:bc], bc, h, N, ld, Wm, bm, gamma, sv["bn"], coef, train, scale,
                                adrop, seed, sid + _CHUNK_STREAM * ci, tf32=prec == ops.PREC_TF32)
            map_gemm(dA, False, kt if bf16 else None, k, dq, b0, bc)
            map_gemm(dA, True, qt if bf16 else None, q, dk, b0, bc)
        del dO, dA
        self._qkv_conv_bwd(P, G, pre, sv, dq, dk, dv, p, B, dxq_acc, dxkv_acc, acc_q, acc_kv)

    def _qkv_conv_bwd(self, P, G, pre, sv, dq, dk, dv, p, B, dxq_acc, dxkv_acc, acc_q=True, acc_kv=True):
        """backward of the three per-patch 3x3 convs (model.py:152-154): data gradients go into dx (accumulated, or written
        when acc_* is False), weight gradients accumulate straight into their slots of the flat gradient buffer"""
        g = self.g
        wq, wk, wv = P[pre + "qconv2d.weight"], P[pre + "kconv2d.weight"], P[pre + "vconv2d.weight"]
        gq, gk, gv = G[pre + "qconv2d.weight"], G[pre + "kconv2d.weight"], G[pre + "vconv2d.weight"]
        xq, xkv = sv["xq"], sv["xkv"]
        C, S = g.C, g.S
        tc = self._prec() == ops.PREC_TF32 and _CONV_MMA["value"]
        if xq is xkv:
            ops.conv3x3_bwd_data([dq, dk, dv], p, [wq, wk, wv], dxq_acc, p, p, B, C, S, S, accumulate=acc_q, tf32=tc)
            ops.conv3x3_bwd_weight(xq, p, [dq, dk, dv], p, [gq, gk, gv], None, p, B, C, S, S, tf32=tc)
        else:
            ops.conv3x3_bwd_data([dq], p, wq, dxq_acc, p, p, B, C, S, S, accumulate=acc_q, tf32=tc)
            ops.conv3x3_bwd_data([dk, dv], p, [wk, wv], dxkv_acc, p, p, B, C, S, S, accumulate=acc_kv, tf32=tc)
            ops.conv3x3_bwd_weight(xq, p, [dq], p, gq, None, p, B, C, S, S, tf32=tc)
            ops.conv3x3_bwd_weight(xkv, p, [dk, dv], p, [gk, gv], None, p, B, C, S, S, tf32=tc)

    def _wgrad(self, dY, X, dW, M, N, K):
        """dW[N,K] += dY[M,N]^T @ X[M,K]; split over the (long) token dimension for parallelism."""
        tiles = ((N + 127) // 128) * ((K + 127) // 128)
        split = max(1, min(64, (2 * 148) // max(tiles, 1), M // 512))
        ops.gemm(dY, X, dW, N, K, M, trans_a=True, trans_b=False, lda=N, ldb=K, ldc=K, accumulate=True,
                 split_k=split, precision=self._prec())

    # ------------------------------------------------------------------------------------------- block
    def _ln_names(self, pre):
        if self.g.shared_ln:
            return pre + "LN.", pre + "LN."
        return pre + "LN1.", pre + "LN2."

    def _block_fwd(self, P, pre, x, l, B, train, seed, sid, saved):
        g = self.g
        N, D, Hd = g.N(l), g.D(l), g.Hd(l)
        n, M = N * D, B * N
        ln1, ln2 = self._ln_names(pre)
        sv_attn = {} if saved is not None else None
        y1 = self._attn_fwd(P, pre + "ReAttn.", x, x, l, B, train, seed, sid, x, sv_attn)
        st1 = _empty((B, 2), x)
        ops.ln_stats(y1, B, n, 1e-5, st1)
        x1 = _empty((B, N, D), x)
        b16 = self._b16()
        # bf16 mode: LayerNorm also writes a bf16 copy of its result (the A operand of the first FeedForward product and
        # of its weight gradient); the fp32 result stays the residual stream.  The hidden tensors are bf16.
        x1b = torch.empty((B, N, D), dtype=torch.bfloat16, device=x.device) if b16 else None
        ops.ln_apply(y1, st1, P[ln1 + "weight"], P[ln1 + "bias"], x1, B, n, out16=x1b)
        ldrop = g.linear_drop if train else 0.0        # Dropout after GELU and after the second Linear (model.py:105,107)
        hdt = torch.bfloat16 if b16 else torch.float32
        pre_act = torch.empty((M, Hd), dtype=hdt, device=x.device) if saved is not None else None
        act = torch.empty((M, Hd), dtype=hdt, device=x.device)
        W1 = self._weight16(P, pre + "FeedForward.net.0.weight") if b16 else P[pre + "FeedForward.net.0.weight"]
        W2 = self._weight16(P, pre + "FeedForward.net.3.weight") if b16 else P[pre + "FeedForward.net.3.weight"]
        self._gemm_tokens(x1b if b16 else x1, W1, act, M, Hd, D,
                          bias=P[pre + "FeedForward.net.0.bias"], act=ops.ACT_GELU, aux_out=pre_act, ldaux=Hd,
                          drop_p=ldrop, drop_seed=seed, drop_stream=sid + 2)
        y2 = _empty((B, N, D), x)
        self._gemm_tokens(act, W2, y2, M, D, Hd,
                          bias=P[pre + "FeedForward.net.3.bias"], residual=x1,
                          drop_p=ldrop, drop_seed=seed, drop_stream=sid + 3)
        st2 = _empty((B, 2), x)
        ops.ln_stats(y2, B, n, 1e-5, st2)
        x2 = _empty((B, N, D), x)
        ops.ln_apply(y2, st2, P[ln2 + "weight"], P[ln2 + "bias"], x2, B, n)
        if saved is not None:
            saved.update(attn=sv_attn, y1=y1, st1=st1, x1=x1b if b16 else x1, pre_act=pre_act, act=act, y2=y2, st2=st2,
                         ldrop=ldrop, seed=seed, sid=sid)
        return x2

    def _block_bwd(self, P, G, pre, dx2, l, B, sv):
        g = self.g
        N, D, Hd = g.N(l), g.D(l), g.Hd(l)
        n, M = N * D, B * N
        prec, b16 = self._prec(), self._b16()
        ln1, ln2 = self._ln_names(pre)
        scratch = _empty((B, ops.LN_SCRATCH), dx2)
        dy2 = torch.empty_like(dx2)
        ldrop, seed, sid = sv["ldrop"], sv["seed"], sv["sid"]
        # bf16 mode: LayerNorm backward also emits the bf16 copy of dy2 that the FeedForward gradient products read
        dy2b = torch.empty(dx2.shape, dtype=torch.bfloat16, device=dx2.device) if (b16 and ldrop == 0) else None
        ops.ln_bwd(dx2, sv["y2"], sv["st2"], P[ln2 + "weight"], dy2, G[ln2 + "weight"], G[ln2 + "bias"], scratch, B, n,
                   dx16=dy2b)
        # FF2: dpre = drop(dy2d @ W2) * gelu'(pre) ; dW2 = dy2d^T act ; db2 = colsum(dy2d), dy2d = drop-mask(dy2)
        if b16:
            dy2d = dy2b if ldrop == 0 else ops.dropout(dy2, torch.empty(dy2.shape, dtype=torch.bfloat16, device=dy2.device),
                                                      ldrop, seed, sid + 3)
            dpre = torch.empty((M, Hd), dtype=torch.bfloat16, device=dx2.device)
            ops.gemm(dy2d, self._weight16(P, pre + "FeedForward.net.3.weight", transposed=True), dpre, M, Hd, D, trans_b=True,
                     lda=D, ldb=D, ldc=Hd, act=ops.ACT_GELU_BWD, aux_in=sv["pre_act"], ldaux=Hd, drop_p=ldrop, drop_seed=seed,
                     drop_stream=sid + 2, precision=prec)
        else:
            W1, W2 = P[pre + "FeedForward.net.0.weight"], P[pre + "FeedForward.net.3.weight"]
            dy2d = ops.dropout(dy2, torch.empty_like(dy2), ldrop, seed, sid + 3) if ldrop > 0 else dy2
            dpre = _empty((M, Hd), dx2)
            ops.gemm(dy2d, W2, dpre, M, Hd, D, trans_b=False, lda=D, ldb=Hd, ldc=Hd, act=ops.ACT_GELU_BWD,
                     aux_in=sv["pre_act"], ldaux=Hd, drop_p=ldrop, drop_seed=seed, drop_stream=sid + 2, precision=prec)
        self._wgrad(dy2d, sv["act"], G[pre + "FeedForward.net.3.weight"], M, D, Hd)
        ops.colsum(dy2d, M, D, D, G[pre + "FeedForward.net.3.bias"], accumulate=True)
        del dy2d, dy2b
        # FF1: dx1 = dpre @ W1 + dy2 ; dW1 = dpre^T x1 ; db1 = colsum(dpre)
        dx1 = torch.empty_like(dx2)
        if b16:
            ops.gemm(dpre, self._weight16(P, pre + "FeedForward.net.0.weight", transposed=True), dx1, M, D, Hd, trans_b=True,
                     lda=Hd, ldb=Hd, ldc=D, residual=dy2, precision=prec)
        else:
            ops.gemm(dpre, W1, dx1, M, D, Hd, trans_b=False, lda=Hd, ldb=D, ldc=D, residual=dy2, precision=prec)
        self._wgrad(dpre, sv["x1"], G[pre + "FeedForward.net.0.weight"], M, Hd, D)
        ops.colsum(dpre, M, Hd, Hd, G[pre + "FeedForward.net.0.bias"], accumulate=True)
        del dy2, dpre
        dy1 = torch.empty_like(dx2)
        ops.ln_bwd(dx1, sv["y1"], sv["st1"], P[ln1 + "weight"], dy1, G[ln1 + "weight"], G[ln1 + "bias"], scratch, B, n)
        del dx1
        # residual: dx = dy1 + (attention path); conv backward accumulates into dy1 in place
        self._attn_bwd(P, G, pre + "ReAttn.", dy1, l, B, sv["attn"], dy1, dy1)
        return dy1

    # ------------------------------------------------------------------------------------------- network
    def forward(self, P: Dict[str, torch.Tensor], X: torch.Tensor, train: bool, save: bool, seed: int = 0):
        g = self.g
        B = X.shape[0]
        C, S = g.C, g.S
        saved: Optional[dict] = {"steps": [], "B": B} if save else None
        self._w16 = {}                  # bf16 weight copies of THIS pass (the optimizer changes the masters between passes)
        self._w16_need_t = bool(save)   # the transposed copies serve the data-gradient products of the backward pass
        src = X
        if g.pe_conv:
            src = torch.empty_like(X)
            ops.conv3x3_fwd(X, 0, P["PE.conv2d.weight"], P["PE.conv2d.bias"], [src], 0, 0, B, C, S, S)
        x = _empty((B, g.N(0), g.D(0)), X)
        ops.pe_fwd(src, 0, P["PE.position_embedding.weight"], g.table_p, x, g.p0, B, C, S, S)
        if save:
            saved["X"] = X
        skips = {}
        sid = 0
        for st in self.sched:
            kind = st[0]
            sv = {} if save else None
            if kind == "block":
                _, pre, l = st
                x = self._block_fwd(P, pre, x, l, B, train, seed, sid, sv)
                sid += 4
            elif kind == "down":
                l = st[1]
                skips[l] = x
                y = _empty((B, g.N(l + 1), g.D(l + 1)), x)
                x = ops.repatch(x, y, B, C, S, S, g.p(l), g.p(l + 1))
            elif kind == "up":
                l = st[1]
                y = _empty((B, g.N(l - 1), g.D(l - 1)), x)
                x = ops.repatch(x, y, B, C, S, S, g.p(l), g.p(l - 1))
            else:
                _, pre, l = st
                enc = skips[l]
                assert enc.shape == x.shape, "enc and dec not same shape"      # model.py:417
                x = self._attn_fwd(P, pre, enc, x, l, B, train, seed, sid, None, sv)
                sid += 2
            if save:
                saved["steps"].append(sv)
        out = _empty((B, C, S, S), X)
        if g.out_conv:
            ops.conv3x3_fwd(x, g.p0, P["conv2d.weight"], P["conv2d.bias"], [out], 0, 0, B, C, S, S)
        else:
            ops.repatch(x, out, B, C, S, S, g.p0, 0)
        if save:
            saved["x_last"] = x
            saved["w16"] = self._w16
        self._w16 = {}
        return out, saved

    def backward(self, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor], saved: dict, dout: torch.Tensor,
                 need_dx: bool):
        """Fills G (zero-initialised gradient tensors keyed like P).  Returns dX or None."""
        g = self.g
        B, C, S = saved["B"], g.C, g.S
        self._w16 = saved.get("w16", {})
        dx = _empty((B, g.N(0), g.D(0)), dout)
        if g.out_conv:
            ops.conv3x3_bwd_data([dout], 0, P["conv2d.weight"], dx, g.p0, 0, B, C, S, S)
            ops.conv3x3_bwd_weight(saved["x_last"], g.p0, [dout], 0, G["conv2d.weight"], G["conv2d.bias"], 0, B, C, S, S)
        else:
            ops.repatch(dout, dx, B, C, S, S, 0, g.p0)
        self._notify("conv2d.")
        skip_grads = {}
        for st, sv in zip(reversed(self.sched), reversed(saved["steps"])):
            kind = st[0]
            if kind == "block":
                _, pre, l = st
                dx = self._block_bwd(P, G, pre, dx, l, B, sv)
                self._notify(pre)
            elif kind == "down":
                l = st[1]
                y = _empty((B, g.N(l), g.D(l)), dx)
                ops.repatch(dx, y, B, C, S, S, g.p(l + 1), g.p(l))
                dx = ops.axpby(skip_grads.pop(l), y, 1.0, 1.0)
            elif kind == "up":
                l = st[1]
                y = _empty((B, g.N(l), g.D(l)), dx)
                dx = ops.repatch(dx, y, B, C, S, S, g.p(l - 1), g.p(l))
            else:
                _, pre, l = st
                d_enc, d_dec = torch.empty_like(dx), torch.empty_like(dx)     # written (not accumulated) by the conv backward
                self._attn_bwd(P, G, pre, dx, l, B, sv, d_enc, d_dec, acc_q=False, acc_kv=False)
                skip_grads[l] = d_enc
                dx = d_dec
                self._notify(pre)
        # PatchEncoder: tokens = patchify(src) + table
        ops.pe_bwd_table(dx, g.p0, G["PE.position_embedding.weight"], g.table_p, B, C, S, S, accumulate=False)
        dX = None
        if g.pe_conv:
            dsrc = _empty((B, C, S, S), dx)
            ops.repatch(dx, dsrc, B, C, S, S, g.p0, 0)
            ops.conv3x3_bwd_weight(saved["X"], 0, [dsrc], 0, G["PE.conv2d.weight"], G["PE.conv2d.bias"], 0, B, C, S, S)
            if need_dx:
                dX = _empty((B, C, S, S), dx)
                ops.conv3x3_bwd_data([dsrc], 0, P["PE.conv2d.weight"], dX, 0, 0, B, C, S, S)
        elif need_dx:
            dX = _empty((B, C, S, S), dx)
            ops.repatch(dx, dX, B, C, S, S, g.p0, 0)
        self._notify("PE.")
        self._w16 = {}
        return dX

    def _notify(self, prefix: str):
        if self.on_grads_ready is not None:
            self.on_grads_ready(prefix)
