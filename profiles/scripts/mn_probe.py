"""Probe how the tensor core reads an MN-major operand tile (debug aid): D = A' . I shows A as the hardware sees it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import Mat, get_runtime
rt = get_runtime("cuda:0")
M, N, K = 128, 32, 32
m = torch.arange(M).view(M, 1).float()
k = torch.arange(K).view(1, K).float()
A = (m * 100 + k).cuda()                 # A[m][k] = 100 m + k
B = torch.eye(N, K).cuda()               # D[m][n] = A[m][n]
for which in ("base", "a_mn", "b_mn"):
    Cd = torch.zeros(M, N, device="cuda")
    if which == "base":
        op = rt.tc_gemm(A=Mat(A.data_ptr(), M, K, K), a_gs=M * K, B=Mat(B.data_ptr(), N, K, K), b_gs=N * K, G=1, passes=3,
                        C=Mat(Cd.data_ptr(), M, N, N), c_gs=M * N, n_tile=32)
    elif which == "a_mn":
        Ast = A.t().contiguous()         # stored [K][M]
        op = rt.tc_gemm(A=Mat(Ast.data_ptr(), K, M, M), a_gs=M * K, B=Mat(B.data_ptr(), N, K, K), b_gs=N * K, G=1, passes=3,
                        C=Mat(Cd.data_ptr(), M, N, N), c_gs=M * N, a_mn=True, n_tile=32)
    else:
        # B' [n][k] = 100 n + k stored [K][N]; A = identity rows -> D[m][n] = B'[n][m] for m < 32
        Bv = (torch.arange(N).view(N, 1).float() * 100 + torch.arange(K).view(1, K).float()).cuda()
        Bst = Bv.t().contiguous()
        Ai = torch.zeros(M, K, device="cuda"); Ai[:K, :K] = torch.eye(K, device="cuda")
        op = rt.tc_gemm(A=Mat(Ai.data_ptr(), M, K, K), a_gs=M * K, B=Mat(Bst.data_ptr(), K, N, N), b_gs=N * K, G=1, passes=3,
                        C=Mat(Cd.data_ptr(), M, N, N), c_gs=M * N, b_mn=True, n_tile=32)
    op(); torch.cuda.synchronize()
    D = Cd.cpu()
    print(which)
    for r in (0, 1, 2, 7, 8, 9, 31, 32, 33):
        print(f"  row {r:3d}:", " ".join(f"{int(v):5d}" for v in D[r, :16].tolist()))
