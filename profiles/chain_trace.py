"""In-kernel timeline of the fused chain kernel (run on the GPU box): per-CTA clock stamps (orlk_tc_set_trace).
Usage: ORLK_PDL=0 python profiles/chain_trace.py"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import GP, get_runtime

rt = get_runtime("cuda:0")
NAMES = ["start", "prev_done", "s0.landed", "s0.mma+partials", "s0.epilogue+store", "s0.cluster_barrier", "s1.strip_issued",
         "s1.prefetch_issued", "s1.thread0_landed", "s1.all_landed", "s1.mma+partials", "s1.epilogue+store", "s1.arrive_issued",
         "s1.cluster_barrier", "s2.all_landed", "s2.mma+partials"]


def trace(G, M, dims, passes):
    gen = torch.Generator().manual_seed(1)
    X = torch.randn(G, M, dims[0], generator=gen).cuda()
    Ws = [(torch.randn(G, dims[i + 1], dims[i], generator=gen) / math.sqrt(dims[i])).cuda() for i in range(len(dims) - 1)]
    bs = [torch.randn(G, dims[i + 1], generator=gen).cuda() for i in range(len(dims) - 1)]
    H = [torch.zeros(G, M, dims[i + 1], device="cuda") for i in range(len(dims) - 1)]
    chains = []
    for g in range(G):
        st = []
        for i in range(len(dims) - 1):
            src = X[g] if i == 0 else H[i - 1][g]
            st.append(GP(A=src.data_ptr(), lda=dims[i], a_layout=0, B=Ws[i][g].data_ptr(), ldb=dims[i], b_layout=1, C=H[i][g].data_ptr(),
                         ldc=dims[i + 1], M=M, N=dims[i + 1], K=dims[i], epi=1, bias=bs[i][g].data_ptr()))
        chains.append(st)
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    L.call("orlk_tc_set_trace", buf.data_ptr())
    op = rt.gemm_chain(chains, passes)
    g = C.c_void_p()
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    for _ in range(3):
        L.call("orlk_graph_launch", g, rt.cur)
    torch.cuda.synchronize()
    L.call("orlk_tc_set_trace", None)
    t = buf.view(-1, 16).cpu()
    n_cta = int((t[:, 0] != 0).sum())
    t = t[:n_cta, :16].double() / 1.9
    rel = t - t[:, 1:2]
    med = rel.median(dim=0).values
    print(f"G={G} M={M} dims={dims} passes={passes} ctas={n_cta}")
    print("   " + "  ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, med.tolist())))
    L.call("orlk_graph_destroy", g)


for passes in (3, 1):
    trace(1, 256, [17, 256, 256, 256, 12], passes)
    trace(2, 256, [23, 256, 256, 256, 1], passes)
