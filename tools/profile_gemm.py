"""Launch a few token-GEMM shapes of the bf16 mode once each under ncu (tools/ncu_gemm.sh): thin / output-bound products
whose in-step rate is far below their HBM roofline, plus the large L0 projection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_unet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cases = [  # (M, N, K, residual, bias)
    (B * 784, 192, 192, False, False), (B * 784, 192, 192, True, True), (B * 196, 768, 64, True, False),
    (B * 784, 192, 32, True, True), (B * 49, 3072, 3072, False, False)]
torch.cuda.profiler.stop() if False else None
bufs = []
for M, N, K, res, bias in cases:
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.empty(M, N, device="cuda")
    R = torch.randn(M, N, device="cuda") if res else None
    bvec = torch.randn(N, device="cuda") if bias else None
    bufs.append((A, W, C, R, bvec, M, N, K))
torch.cuda.synchronize()
for rep in range(2):
    if rep == 1:
        torch.cuda.cudart().cudaProfilerStart()
    for A, W, C, R, bvec, M, N, K in bufs:
        ops.gemm(A, W, C, M, N, K, trans_b=True, lda=K, ldb=K, ldc=N, bias=bvec, residual=R, precision=ops.PREC_TF32)
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
