"""In-kernel timeline of the small-row GEMM (run on the GPU box): per-CTA clock stamps (orlk_tc_set_trace).
Usage: ORLK_PDL=0 python profiles/tiny_trace.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import GP, get_runtime

rt = get_runtime("cuda:0")
NAMES = ["start", "prev_done", "issued", "landed", "k_loop", "reduced", "stored"]


def trace(M, N, K, a_layout, b_layout, G=1, epi=0, passes=0):
    keep, probs = [], []
    for g in range(G):
        A = torch.randn(M, K, device="cuda") if a_layout == 0 else torch.randn(K, M, device="cuda")
        B = torch.randn(K, N, device="cuda") if b_layout == 0 else torch.randn(N, K, device="cuda")
        Cd = torch.zeros(M, N, device="cuda")
        CT = torch.zeros(N, M, device="cuda")
        bias = torch.randn(N, device="cuda")
        keep += [A, B, Cd, CT, bias]
        probs.append(GP(A=A.data_ptr(), lda=A.stride(0), a_layout=a_layout, B=B.data_ptr(), ldb=B.stride(0), b_layout=b_layout,
                        C=Cd.data_ptr(), ldc=N, CT=CT.data_ptr(), ldct=M, M=M, N=N, K=K, epi=epi, bias=bias.data_ptr()))
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    L.call("orlk_tc_set_trace", buf.data_ptr())
    op = rt.gemm(probs, L.CFG_TINY, passes=passes)
    g = C.c_void_p()
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    L.call("orlk_tc_set_trace", None)
    for _ in range(3):
        L.call("orlk_graph_launch", g, rt.cur)
    torch.cuda.synchronize()
    t = buf.view(-1, 16).cpu()
    n_cta = int((t[:, 0] != 0).sum())
    t = t[:n_cta, :7].double() / 1.9
    rel = t - t[:, 1:2]
    med = rel.median(dim=0).values
    mx = rel.max(dim=0).values
    print(f"M={M} N={N} K={K} a{a_layout} b{b_layout} G={G} passes={passes} ctas={n_cta}")
    print("   median " + "  ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, med.tolist())))
    print("   max    " + "  ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, mx.tolist())))
    L.call("orlk_graph_destroy", g)


for passes in (0, 3):
    trace(256, 256, 256, 0, 1, passes=passes)          # forward (A k-contiguous, W [out][in])
    trace(256, 256, 256, 0, 0, passes=passes)          # dgrad (W [k][n])
    trace(256, 256, 256, 0, 1, G=2, passes=passes)
    trace(256, 256, 4, 0, 1, passes=passes)
