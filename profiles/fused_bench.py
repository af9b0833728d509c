"""Per-launch device time of the fused critic passes alone (run on the GPU box): R repeats captured in one graph,
timed with CUDA events.  Usage: python profiles/fused_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import Mat, get_runtime

rt = get_runtime("cuda:0")


def al(n):
    return (n + 3) // 4 * 4


def time_graph(op, reps=20, iters=20):
    g = C.c_void_p()
    torch.cuda.synchronize()
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    for _ in range(reps):
        op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    for _ in range(3):
        L.call("orlk_graph_launch", g, rt.cur)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.call("orlk_graph_launch", g, rt.cur)
    e1.record()
    torch.cuda.synchronize()
    L.call("orlk_graph_destroy", g)
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def fwd_op(M, N, K0, nh, G):
    ldx = al(K0)
    X = torch.randn(M, ldx, device="cuda")
    dims = [(N, K0)] + [(N, N)] * (nh - 1) + [(1, N)]
    offs, off = [], 0
    for (o, i) in dims:
        w = off; off = al(off + o * i)
        b = off; off = al(off + o)
        offs.append((w, b))
    block = al(off)
    P = torch.randn(G * block, device="cuda") / 16
    Plo = torch.zeros_like(P)
    pad = torch.zeros(2, G, N, 32, device="cuda")
    rt.fused_prep(P, Plo, W0=P.data_ptr() + 4 * offs[0][0], gs=block, N=N, K0=K0, G=G, w0pad=pad)()
    H = [torch.zeros(G, M, N, device="cuda") for _ in range(nh)]
    out = torch.zeros(G, M, device="cuda")
    base, lo = P.data_ptr(), Plo.data_ptr()
    job = rt.fused_fwd_job(X=Mat(X.data_ptr(), M, K0, ldx), W0pad=pad[0].data_ptr(), W0pad_lo=pad[1].data_ptr(),
                           W=[0] + [base + 4 * offs[l][0] for l in range(1, nh)],
                           Wlo=[0] + [lo + 4 * offs[l][0] for l in range(1, nh)], bias=[base + 4 * offs[l][1] for l in range(nh)],
                           H=[h.data_ptr() for h in H], gs=block, h_gs=M * N, head_w=base + 4 * offs[nh][0],
                           head_b=base + 4 * offs[nh][1], out=out.data_ptr(), out_gs=M, M=M, N=N, K0=K0, G=G)
    op = rt.critic_fwd_fused([job])
    keep = (X, P, Plo, H, out, pad)
    return lambda op=op, keep=keep: op()


if __name__ == "__main__":
    tag = "no_store" if os.environ.get("ORLK_FUSED_NO_STORE") == "1" else "normal"
    for (M, G) in ((7936, 2), (128, 1), (3968, 2)):
        op = fwd_op(M, 256, 23, 3, G)
        print(f"fwd fused {tag}: M={M} G={G}: {time_graph(op):.2f} us per launch", flush=True)
