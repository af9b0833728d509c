"""Per-launch time of the small-row GEMM in a graph of 20 dependent repeats (run on the GPU box)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from offlinerlkit_b200 import _lib as L
from offlinerlkit_b200.engine.core import GP, get_runtime
rt = get_runtime("cuda:0")

def timeit(op, reps=20):
    e0, e1 = C.c_void_p(), C.c_void_p()
    L.call("orlk_event_create", C.byref(e0)); L.call("orlk_event_create", C.byref(e1))
    g = C.c_void_p()
    torch.cuda.synchronize()
    rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
    L.call("orlk_graph_begin", rt.cur)
    for _ in range(reps): op()
    L.call("orlk_graph_end", rt.cur, C.byref(g))
    rt.cur = rt.exec_ptr
    for _ in range(3): L.call("orlk_graph_launch", g, rt.cur)
    L.call("orlk_event_record", e0, rt.cur)
    for _ in range(5): L.call("orlk_graph_launch", g, rt.cur)
    L.call("orlk_event_record", e1, rt.cur)
    ms = C.c_float(); L.call("orlk_event_elapsed_ms", e0, e1, C.byref(ms))
    return 1e3 * ms.value / reps / 5

def prob(M, N, K, G=1):
    keep, probs = [], []
    for g in range(G):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); Cd = torch.zeros(M, N, device="cuda")
        bias = torch.randn(N, device="cuda")
        keep += [A, B, Cd, bias]
        probs.append(GP(A=A.data_ptr(), lda=K, a_layout=0, B=B.data_ptr(), ldb=K, b_layout=1, C=Cd.data_ptr(), ldc=N, M=M, N=N, K=K,
                        epi=1, bias=bias.data_ptr()))
    return probs, keep

for (M, N, K, G) in [(256, 256, 256, 1), (256, 256, 256, 2), (512, 256, 256, 1), (256, 256, 4, 1), (32, 16, 4, 1), (256, 256, 17, 1)]:
    probs, keep = prob(M, N, K, G)
    for passes in (0, 3):
        op = rt.gemm(probs, L.CFG_TINY, passes=passes)
        print(f"M={M} N={N} K={K} G={G} passes={passes}: {timeit(op):6.2f} us/launch")
e = torch.zeros(4, device="cuda")
print("philox (tiny elementwise):", timeit(lambda: L.call("orlk_philox_fill", e.data_ptr(), 1, 0, 0.0, 1.0, 1, None, None, rt.cur)))
