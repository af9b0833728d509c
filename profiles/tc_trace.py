"""In-kernel timeline of the tcgen05 GEMM (run on the GPU box): per-CTA %globaltimer stamps written by the kernel
itself (orlk_tc_set_trace).  Prints, per shape, the median offset in ns of each phase from the CTA's start and the
span between the first CTA start and the last CTA end.  Usage: python profiles/tc_trace.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import Mat, get_runtime

rt = get_runtime("cuda:0")
NAMES = ["start", "prev_done", "prologue", "tile0", "mma0", "mmaN", "accum", "stores"] + [f"slab{i}" for i in range(8)]


def trace(G, M, N, K, passes, n_tile, reps=1, k_splits=1, out=True):
    A = torch.randn(G, M, K, device="cuda")
    B = torch.randn(G, N, K, device="cuda")
    Cb = torch.zeros(max(k_splits, 1), G, M, N, device="cuda")
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
    kw = dict(C=Mat(Cb.data_ptr(), M, N, N), c_gs=M * N, c_split_stride=G * M * N) if out else {}
    op = rt.tc_gemm(A=Mat(A.data_ptr(), M, K, K), a_gs=M * K, B=Mat(B.data_ptr(), N, K, K), b_gs=N * K, G=G, passes=passes,
                    n_tile=n_tile, k_splits=k_splits, **kw)
    g = C.c_void_p()
    torch.cuda.synchronize()
    L.call("orlk_tc_set_trace", buf.data_ptr())
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    for _ in range(reps):
        op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    L.call("orlk_tc_set_trace", None)
    for _ in range(3):
        L.call("orlk_graph_launch", g, rt.cur)
    torch.cuda.synchronize()
    t = buf.view(-1, 16).cpu()
    n_cta = int((t[:, 0] != 0).sum())
    t = t[:n_cta, :16].double() / 1.9        # SM clocks -> ~ns
    rel = t - t[:, 1:2]                      # relative to "predecessor complete"
    rel[t == 0] = float("nan")
    med = rel.nanmedian(dim=0).values
    span = float(t[:, 7].max() - t[:, 0].min())
    skew = float(t[:, 0].max() - t[:, 0].min())
    print(f"G={G} M={M} N={N} K={K} p{passes} nt={n_tile} splits={k_splits} ctas={n_cta}: span {span:.0f} ns, start skew {skew:.0f} ns")
    print("   " + "  ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, med.tolist()) if v == v))
    L.call("orlk_graph_destroy", g)


for passes in (1, 3):
    trace(1, 256, 256, 256, passes, 32)      # 16 CTAs, 8 share each A tile
    trace(1, 128, 32, 256, passes, 32)       # 1 CTA alone, 20 KB per slab
    trace(1, 128, 256, 256, passes, 256)     # 1 CTA alone, 48 KB per slab
    trace(2, 7936, 256, 256, passes, 256)
