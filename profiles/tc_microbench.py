"""Micro-benchmark of the tcgen05 GEMM launcher (run on the GPU box): device time per launch for the critic-pass
shapes, by precision mode and fused-output set.  Usage: python profiles/tc_microbench.py [out.json]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import Mat, get_runtime

rt = get_runtime("cuda:0")
G, M, N, K = 2, 7936, 256, 256
A = torch.randn(G, M, K, device="cuda")
B = torch.randn(G, N, K, device="cuda")
Cb = torch.zeros(16, G, M, N, device="cuda")
CT = torch.zeros(G, N, M, device="cuda")
bias = torch.randn(G, N, device="cuda")
aux = torch.randn(G, M, N, device="cuda")
AT = torch.randn(G, N, M, device="cuda")      # wgrad operands [out, M]
BT = torch.randn(G, N, M, device="cuda")
Wg = torch.zeros(16, G, N, N, device="cuda")
rs = torch.zeros(16, G, N, device="cuda")


def timeit(op, reps=30):
    e0, e1 = C.c_void_p(), C.c_void_p()
    L.call("orlk_event_create", C.byref(e0))
    L.call("orlk_event_create", C.byref(e1))
    g = C.c_void_p()
    torch.cuda.synchronize()
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    for _ in range(reps):
        op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    for _ in range(2):
        L.call("orlk_graph_launch", g, rt.cur)
    L.call("orlk_event_record", e0, rt.cur)
    L.call("orlk_graph_launch", g, rt.cur)
    L.call("orlk_event_record", e1, rt.cur)
    ms = C.c_float()
    L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
    return 1e3 * ms.value / reps


out = {}
for passes in (1, 3):
    for name, kw in {
        "fwd none": dict(),
        "fwd C": dict(C=Mat(Cb.data_ptr(), M, N, N), c_gs=M * N),
        "fwd CT": dict(CT=Mat(CT.data_ptr(), N, M, M), ct_gs=N * M),
        "fwd C+CT+bias+relu": dict(C=Mat(Cb.data_ptr(), M, N, N), c_gs=M * N, CT=Mat(CT.data_ptr(), N, M, M), ct_gs=N * M,
                                   bias=bias.data_ptr(), bias_gs=N, epi=L.EPI_RELU),
        "dgrad C+CT+mask": dict(C=Mat(Cb.data_ptr(), M, N, N), c_gs=M * N, CT=Mat(CT.data_ptr(), N, M, M), ct_gs=N * M,
                                aux=Mat(aux.data_ptr(), M, N, N), aux_gs=M * N, epi=L.EPI_RELU_MASK),
    }.items():
        op = rt.tc_gemm(A=Mat(A.data_ptr(), M, K, K), a_gs=M * K, B=Mat(B.data_ptr(), N, K, K), b_gs=N * K, G=G,
                        passes=passes, **kw)
        out[f"p{passes} {name}"] = timeit(op)
    for splits in (8, 16, 31):
        op = rt.tc_gemm(A=Mat(AT.data_ptr(), N, M, M), a_gs=N * M, B=Mat(BT.data_ptr(), N, M, M), b_gs=N * M, G=G,
                        passes=passes, C=Mat(Wg.data_ptr(), N, N, N), c_gs=N * N, c_split_stride=G * N * N,
                        rowsum=rs.data_ptr(), rowsum_gs=N, rowsum_split_stride=G * N, k_splits=splits)
        out[f"p{passes} wgrad splits{splits}"] = timeit(op)
for k, v in out.items():
    print(f"{k:32s} {v:8.2f} us")
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)

# ---- small-M study: where do ~10 us per layer go?
print("small-M (G=2, M=256, N=256, n_tile=32):")
As = torch.randn(2, 256, 256, device="cuda")
Bs = torch.randn(2, 256, 256, device="cuda")
Cs = torch.zeros(2, 256, 256, device="cuda")
for passes in (1, 3):
    for K in (32, 256):
        for nt in (32, 64, 256):
            for outs in ("none", "C"):
                kw = dict(C=Mat(Cs.data_ptr(), 256, 256, 256), c_gs=256 * 256) if outs == "C" else {}
                op = rt.tc_gemm(A=Mat(As.data_ptr(), 256, K, 256), a_gs=256 * 256, B=Mat(Bs.data_ptr(), 256, K, 256), b_gs=256 * 256,
                                G=2, passes=passes, n_tile=nt, **kw)
                print(f"  p{passes} K={K:3d} n_tile={nt:3d} out={outs:4s} {timeit(op):7.2f} us")
# an empty-ish kernel for the launch floor
e = torch.zeros(1, device="cuda")
print("  launch floor (philox 4 elements):", timeit(lambda: L.call("orlk_philox_fill", e.data_ptr(), 1, 0, 0.0, 1.0, 1, None, None, rt.cur)))
