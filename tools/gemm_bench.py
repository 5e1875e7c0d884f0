"""Micro-benchmark of vu_gemm (tcgen05 TF32 path) on the token-GEMM shapes of a Base training step at B images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_unet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
BF16 = len(sys.argv) > 2 and sys.argv[2] == "bf16"      # bf16 mode: bf16 operands, K-major transposed weight copy for dgrad
prec = ops.PREC_TF32
shapes = []
for name, N_tok, D, Hd in (("L0", 49, 3072, 128), ("L1", 196, 768, 64), ("L2", 784, 192, 32)):
    M = B * N_tok
    shapes += [(f"{name} proj fwd  NT", M, D, D, False, True), (f"{name} proj dgrad NN", M, D, D, False, BF16),
               (f"{name} proj wgrad TN", D, D, M, True, False), (f"{name} FF1 fwd NT", M, Hd, D, False, True),
               (f"{name} FF2 fwd NT", M, D, Hd, False, True), (f"{name} FF1 wgrad TN", Hd, D, M, True, False)]
tot_ms = 0
for name, M, N, K, ta, tb in shapes:
    A = torch.randn((K, M) if ta else (M, K), device="cuda")
    Bm = torch.randn((N, K) if tb else (K, N), device="cuda")
    if BF16:
        A, Bm = A.bfloat16(), Bm.bfloat16()
    C = torch.zeros(M, N, device="cuda")
    split = 1
    if ta:
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        split = max(1, min(64, (2 * 148) // max(tiles, 1), K // 512))
    def run():
        ops.gemm(A, Bm, C, M, N, K, trans_a=ta, trans_b=tb, lda=A.shape[1], ldb=Bm.shape[1], ldc=N,
                 accumulate=ta, split_k=split, precision=prec)
    for _ in range(3): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot_ms += ms
    print(f"{name:22s} M={M:6d} N={N:5d} K={K:6d} split={split:2d}  {ms*1e3:8.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s")
print("sum", round(tot_ms, 3), "ms")
