#!/usr/bin/env python
"""Time the scores GEMM (S = Q K^T, dA = dO V^T) at the Base bottleneck shape against a plain device fill."""
import sys, torch
sys.path.insert(0, ".")
from vit_unet_b200 import ops
B, h = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 8
N = int(sys.argv[3]) if len(sys.argv) > 3 else 784
hd = int(sys.argv[2]) if len(sys.argv) > 2 else 24
D = h * hd
q = torch.randn(B, N, D, device="cuda"); k = torch.randn(B, N, D, device="cuda")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for dt in (torch.float32, torch.bfloat16):
    S = torch.empty(B, h, N, N, dtype=dt, device="cuda")
    f = lambda: ops.gemm(q, k, S, N, N, hd, trans_b=True, lda=D, ldb=D, ldc=N, batch_outer=B, batch_inner=h,
                         sA=(N * D, hd), sB=(N * D, hd), sC=(h * N * N, N * N), precision=ops.PREC_TF32)
    ms = timeit(f); gb = S.numel() * S.element_size() / 1e9
    ms_fill = timeit(lambda: S.fill_(1.0))
    print(f"{dt}: scores {ms:.3f} ms = {gb/ms*1e3:.0f} GB/s   fill {ms_fill:.3f} ms = {gb/ms_fill*1e3:.0f} GB/s")
    del S
