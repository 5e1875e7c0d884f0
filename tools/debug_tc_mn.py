import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vit_unet_b200 import ops
M, N, K = 128, 128, 32
torch.set_printoptions(linewidth=200, sci_mode=False)
for krow in (0, 1, 9, 17):
    A = torch.zeros(M, K); A[:, krow] = 1.0
    for what in ("n", "k"):
        Bm = torch.zeros(K, N)
        for k in range(K):
            for n in range(N):
                Bm[k, n] = n if what == "n" else k
        out = torch.full((M, N), -1.0, device="cuda")
        ops.gemm(A.cuda(), Bm.cuda(), out, M, N, K, trans_a=False, trans_b=False, lda=K, ldb=N, ldc=N, precision=ops.PREC_TF32)
        torch.cuda.synchronize()
        o = out.cpu()
        print(f"A one-hot k={krow}, B={what}: row0 cols 0..40:", o[0, :40].int().tolist())
        print("   cols 64..72:", o[0, 64:72].int().tolist(), " row 77 cols 0..8:", o[77, :8].int().tolist())
