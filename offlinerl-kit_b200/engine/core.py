"""Core host plumbing: device views, GEMM/Adam descriptor tables, launch plans (CUDA graph capture)."""
import ctypes as C
import gc
import os
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib as L

NUM_SMS = 148
_ctypes_pointer = C.pointer      # (the name ``C`` is shadowed by a matrix argument in Runtime.tc_gemm)


def require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise L.OrlkError(f"offlinerlkit_b200 needs a CUDA device (got {dev}); there is no CPU fallback")
    L.load()
    return dev


@dataclass
class Mat:
    """A row-major fp32 device matrix view: element (r,c) at ptr + 4*(r*ld + c)."""
    ptr: int
    rows: int
    cols: int
    ld: int
    keep: object = None     # tensor that owns the memory (kept alive)

    @staticmethod
    def of(t: torch.Tensor) -> "Mat":
        assert t.dtype == torch.float32 and t.is_cuda
        if t.dim() == 1:
            t = t.view(1, -1)
        assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
        return Mat(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0) if t.shape[0] > 1 else t.shape[1], t)

    def rows_(self, r0: int, r1: int) -> "Mat":
        assert 0 <= r0 <= r1 <= self.rows
        return Mat(self.ptr + 4 * r0 * self.ld, r1 - r0, self.cols, self.ld, self.keep)

    def cols_(self, c0: int, c1: int) -> "Mat":
        assert 0 <= c0 <= c1 <= self.cols
        return Mat(self.ptr + 4 * c0, self.rows, c1 - c0, self.ld, self.keep)


@dataclass
class GP:
    """One GEMM problem C[M,N] = epi(A x B) in the terms of include/orlk_b200.h:OrlkGemmDesc."""
    A: int
    lda: int
    a_layout: int
    B: int
    ldb: int
    b_layout: int
    C: int
    ldc: int
    M: int
    N: int
    K: int
    epi: int = L.EPI_NONE
    bias: int = 0
    aux: int = 0
    ldaux: int = 0
    C2: int = 0
    rowsum: int = 0
    colsum: int = 0
    k_splits: int = 1
    split_base: int = 0
    c_split_stride: int = 0
    sum_split_stride: int = 0
    CT: int = 0
    ldct: int = 0


class Runtime:
    """One CUDA stream + the loaded library + a keep-alive list for descriptor tables."""

    def __init__(self, device):
        self.device = require_cuda(device)
        self.lib = L.load()
        torch.cuda.set_device(self.device)
        # Launches go to torch's current stream (normally the legacy default stream) so that they are ordered with
        # the caller's own torch work; graph capture needs a non-default stream, so Plan.capture() temporarily
        # switches `cur` to a private one.
        self.capture_stream = torch.cuda.Stream(device=self.device)
        self.exec_ptr = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        self.cur = self.exec_ptr
        self._keep: List[torch.Tensor] = []
        L.call("orlk_tc_init")
        L.call("orlk_fused_init")
        L.call("orlk_gemm_init")
        L.call("orlk_gemm_tiny_init")
        L.call("orlk_gemm_chain_init")
        L.call("orlk_narrow_init")

    # ---- memory helpers (torch owns device memory: plumbing)
    def zeros(self, *shape, dtype=torch.float32) -> torch.Tensor:
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def upload_bytes(self, raw: bytes) -> torch.Tensor:
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self._keep.append(t)
        return t

    def sync(self) -> None:
        L.call("orlk_stream_sync", self.cur)

    # ---- launch builders: each returns a zero-argument closure that enqueues on self.stream
    def gemm(self, problems: Sequence[GP], cfg: int, passes: int = 0) -> Callable[[], None]:
        """Grouped fp32 GEMM launch(es): the kernel is specialised on the operand layouts, so problems are bucketed
        by (a_layout, b_layout) -- one launch per bucket (normally a single one).  ``passes`` (CFG_TINY only):
        0 = fp32 FFMA, 3 = 3xTF32 tensor-core MMAs (fp32-grade), 1 = single-pass TF32."""
        buckets: Dict[Tuple[int, int], List[GP]] = {}
        for p in problems:
            buckets.setdefault((p.a_layout, p.b_layout), []).append(p)
        ops = [self._gemm_bucket(ps, cfg, key, passes) for key, ps in buckets.items()]
        if len(ops) == 1:
            return ops[0]
        return lambda: [op() for op in ops] and None

    def _gemm_bucket(self, problems: Sequence[GP], cfg: int, layouts: Tuple[int, int], passes: int = 0) -> Callable[[], None]:
        BM, BN, BK = L.CFG_TILES[cfg]
        arr = (L.GemmDesc * len(problems))()
        tile = 0
        for i, p in enumerate(problems):
            d = arr[i]
            tiles_m, tiles_n = -(-p.M // BM), -(-p.N // BN)
            splits = max(1, p.k_splits)
            chunk = -(-p.K // splits)
            chunk = -(-chunk // BK) * BK
            splits = max(1, -(-p.K // chunk))
            if p.epi != L.EPI_NONE or p.bias:
                assert splits == 1, "epilogues need the full k range"
            d.A, d.B, d.C, d.C2 = p.A, p.B, p.C, p.C2 or None
            d.bias, d.aux, d.rowsum, d.colsum = p.bias or None, p.aux or None, p.rowsum or None, p.colsum or None
            d.lda, d.ldb, d.ldc, d.ldaux = p.lda, p.ldb, p.ldc, p.ldaux
            d.CT, d.ldct = p.CT or None, p.ldct
            d.c_split_stride, d.sum_split_stride = p.c_split_stride, p.sum_split_stride
            d.M, d.N, d.K = p.M, p.N, p.K
            d.a_layout, d.b_layout, d.epi = p.a_layout, p.b_layout, p.epi
            d.k_splits, d.k_chunk, d.split_base = splits, chunk, p.split_base
            d.tile_start, d.tiles_m, d.tiles_n = tile, tiles_m, tiles_n
            tile += tiles_m * tiles_n * splits
        al, bl = layouts
        if cfg == L.CFG_TINY:
            # the small-row kernel takes its problems by value in the kernel parameters, at most 16 per launch
            ops = []
            for i0 in range(0, len(problems), 16):
                n = min(16, len(problems) - i0)
                sub = (L.GemmDesc * n)()
                base = arr[i0].tile_start
                for i in range(n):
                    C.memmove(C.byref(sub[i]), C.byref(arr[i0 + i]), C.sizeof(L.GemmDesc))
                    assert sub[i].k_splits == 1, "the small-row kernel does not split k"
                    sub[i].tile_start -= base
                tiles = (arr[i0 + n].tile_start if i0 + n < len(problems) else tile) - base
                ops.append(lambda sub=sub, n=n, tiles=tiles: L.call("orlk_gemm_tiny", sub, n, tiles, al, bl, passes, self.cur))
            return ops[0] if len(ops) == 1 else (lambda: [op() for op in ops] and None)
        dev = self.upload_bytes(bytes(arr))
        n, total, ptr = len(problems), tile, C.c_void_p(dev.data_ptr())
        return lambda: L.call("orlk_gemm_grouped", ptr, n, total, cfg, al, bl, self.cur)

    def gemm_chain(self, chains: Sequence[Sequence[GP]], passes: int, passes0: Optional[int] = None) -> Callable[[], None]:
        """One launch for whole small-row layer chains (csrc/orlk_chain.cu): ``chains[c][s]`` is stage s of chain c, the
        A operand of stage s+1 is the C output of stage s; all chains have the same number of stages and rows."""
        n_chains, n_stages = len(chains), len(chains[0])
        assert all(len(c) == n_stages for c in chains) and n_chains * n_stages <= 24
        arr = (L.GemmDesc * (n_chains * n_stages))()
        for c, chain in enumerate(chains):
            for s_, p in enumerate(chain):
                d = arr[c * n_stages + s_]
                assert p.a_layout == 0 and p.k_splits <= 1 and not p.rowsum and not p.colsum and not p.CT, p
                assert s_ == 0 or (p.A == chain[s_ - 1].C and p.K == chain[s_ - 1].N), "stage input must be the previous output"
                d.A, d.B, d.C, d.C2 = p.A, p.B, p.C, p.C2 or None
                d.bias, d.aux = p.bias or None, p.aux or None
                d.lda, d.ldb, d.ldc, d.ldaux = p.lda, p.ldb, p.ldc, p.ldaux
                d.M, d.N, d.K = p.M, p.N, p.K
                d.a_layout, d.b_layout, d.epi = p.a_layout, p.b_layout, p.epi
                d.k_splits, d.k_chunk = 1, p.K
        p0 = passes if passes0 is None else passes0
        return lambda: L.call("orlk_gemm_chain", arr, n_chains, n_stages, passes, p0, self.cur)

    def tc_gemm(self, *, A: Mat, a_gs: int, B: Mat, b_gs: int, G: int, passes: int, epi: int = L.EPI_NONE,
                C: Optional[Mat] = None, c_gs: int = 0, c_split_stride: int = 0, CT: Optional[Mat] = None, ct_gs: int = 0,
                bias: int = 0, bias_gs: int = 0, aux: Optional[Mat] = None, aux_gs: int = 0, rowsum: int = 0,
                rowsum_gs: int = 0, rowsum_split_stride: int = 0, k_splits: int = 1, n_tile: int = 0,
                gen_row: int = 0, gen_row_gs: int = 0, gen_col: int = 0, gen_col_gs: int = 0,
                a_mn: bool = False, b_mn: bool = False) -> Callable[[], None]:
        """tcgen05 GEMM launch: C[g] = epi(A[g] (M x K) . B[g]^T (N x K)); group strides in floats.
        ``a_mn`` / ``b_mn``: the Mat describes the operand as STORED, [K rows][M or N columns]."""
        q = L.TcGemm()
        q.A, q.lda, q.a_gs = A.ptr, A.ld, a_gs
        q.B, q.ldb, q.b_gs = B.ptr, B.ld, b_gs
        if C is not None:
            q.C, q.ldc, q.c_gs, q.c_split_stride = C.ptr, C.ld, c_gs, c_split_stride
        if CT is not None:
            q.CT, q.ldct, q.ct_gs = CT.ptr, CT.ld, ct_gs
        q.bias, q.bias_gs = bias or None, bias_gs
        if aux is not None:
            q.aux, q.ldaux, q.aux_gs = aux.ptr, aux.ld, aux_gs
        q.rowsum, q.rowsum_gs, q.rowsum_split_stride = rowsum or None, rowsum_gs, rowsum_split_stride
        q.a_mn, q.b_mn = int(a_mn), int(b_mn)
        Ma, Ka = (A.cols, A.rows) if a_mn else (A.rows, A.cols)
        Nb, Kb = (B.cols, B.rows) if b_mn else (B.rows, B.cols)
        assert Ka == Kb, (Ka, Kb)
        q.M, q.N, q.K, q.G = Ma, Nb, Ka, G
        q.epi, q.passes, q.n_tile = epi, passes, n_tile
        q.gen_row, q.gen_row_gs, q.gen_col, q.gen_col_gs = gen_row or None, gen_row_gs, gen_col or None, gen_col_gs
        q.k_splits = self.lib.orlk_tc_effective_splits(Ka, k_splits)
        qp = _ctypes_pointer(q)      # the struct is read on the host at every launch: keep it alive in the closure
        return lambda: L.call("orlk_tc_gemm", qp, self.cur)

    @staticmethod
    def fused_fwd_job(*, X: Mat, W0pad: int, W0pad_lo: int, W: Sequence[int], Wlo: Sequence[int], bias: Sequence[int],
                      H: Optional[Sequence[int]], gs: int, h_gs: int, head_w: int, head_b: int, out: int, out_gs: int,
                      M: int, N: int, K0: int, G: int, relu_bits: int = 0) -> dict:
        """One pass of ``critic_fwd_fused``: W / Wlo / bias / H are per hidden layer (W[0], Wlo[0] unused: the first layer
        reads the padded copies); H = None keeps no activations (target / inference pass)."""
        return dict(X=X, W0pad=W0pad, W0pad_lo=W0pad_lo, W=list(W), Wlo=list(Wlo), bias=list(bias),
                    H=list(H) if H is not None else None, gs=gs, h_gs=h_gs, head_w=head_w, head_b=head_b, out=out,
                    out_gs=out_gs, M=M, N=N, K0=K0, G=G, relu_bits=relu_bits)

    def critic_fwd_fused(self, jobs: Sequence[dict], pairs: Optional[bool] = None) -> Callable[[], None]:
        """Whole Linear+ReLU critic passes + scalar heads for all members in ONE tcgen05 launch (csrc/orlk_fused.cu);
        one or two jobs (``fused_fwd_job``) side by side.  ``pairs``: CTA pairs (cta_group::2); None = when the launch
        has at least 32 strips (ORLK_FUSED_2CTA=0 / 1 forces it off / on)."""
        assert 1 <= len(jobs) <= 2
        if pairs is None:
            env = os.environ.get("ORLK_FUSED_2CTA", "auto")
            strips = sum(j["G"] * (-(-j["M"] // 128)) for j in jobs)
            pairs = env == "1" or (env not in ("0", "1") and strips >= 32)
        arr = (L.FusedFwd * len(jobs))()
        for q, j in zip(arr, jobs):
            q.X, q.ldx = j["X"].ptr, j["X"].ld
            q.W0pad, q.W0pad_lo = j["W0pad"], j["W0pad_lo"]
            for l in range(len(j["bias"])):
                q.W[l], q.Wlo[l], q.bias[l] = (j["W"][l] or None), (j["Wlo"][l] or None), j["bias"][l]
                q.H[l] = j["H"][l] if j["H"] is not None else None
            q.gs, q.h_gs = j["gs"], j["h_gs"]
            q.head_w, q.head_b, q.out, q.out_gs = j["head_w"], j["head_b"], j["out"], j["out_gs"]
            q.relu_bits = j.get("relu_bits") or None
            q.M, q.N, q.K0, q.G, q.n_hidden = j["M"], j["N"], j["K0"], j["G"], len(j["bias"])
            q.flags = L.FUSED_PAIRS if pairs else 0
        n = len(jobs)
        return lambda: L.call("orlk_critic_fwd_fused", arr, n, self.cur)

    def critic_bwd_fused(self, *, dq: int, dq_gs: int, head_w: int, relu_bits: int, WT: Sequence[int], WTlo: Sequence[int],
                         dZ: Sequence[int], gs: int, dz_gs: int, M: int, N: int, G: int,
                         pairs: Optional[bool] = None) -> Callable[[], None]:
        """Input-gradient chain of a fused critic pass behind its scalar head, one tcgen05 launch (csrc/orlk_fused.cu).
        WT / WTlo per hidden layer ([0] unused), dZ[l] for l = 0 .. n_hidden-2."""
        q = L.FusedBwd()
        q.dq, q.dq_gs, q.head_w, q.relu_bits = dq, dq_gs, head_w, relu_bits
        for l in range(len(WT)):
            q.WT[l], q.WTlo[l] = (WT[l] or None), (WTlo[l] or None)
        for l in range(len(dZ)):
            q.dZ[l] = dZ[l]
        q.gs, q.dz_gs, q.M, q.N, q.G, q.n_hidden = gs, dz_gs, M, N, G, len(WT)
        if pairs is None:       # CTA pairs (cta_group::2) from 32 strips on, as for the forward pass
            env = os.environ.get("ORLK_FUSED_2CTA", "auto")
            pairs = env == "1" or (env not in ("0", "1") and G * (-(-M // 128)) >= 32)
        q.flags = L.FUSED_PAIRS if pairs else 0
        qp = _ctypes_pointer(q)
        return lambda: L.call("orlk_critic_bwd_fused", qp, self.cur)

    def fused_prep_multi(self, jobs: Sequence[dict]) -> Callable[[], None]:
        """``fused_prep`` for up to four arenas in one launch; jobs: dicts with the keyword arguments of ``fused_prep``."""
        assert 1 <= len(jobs) <= 4
        arr = (L.FusedPrep * len(jobs))()
        keep = []
        for q, j in zip(arr, jobs):
            q.src, q.dst_lo, q.n = j["src"].data_ptr(), j["dst_lo"].data_ptr(), j["src"].numel()
            q.W0, q.gs = (j.get("W0") or None), j.get("gs", 0)
            q.w0pad = j["w0pad"].data_ptr() if j.get("w0pad") is not None else None
            q.N, q.K0, q.G = j.get("N", 0), j.get("K0", 0), j.get("G", 0)
            keep.append((j["src"], j["dst_lo"], j.get("w0pad")))
        n = len(jobs)
        return lambda keep=keep: L.call("orlk_fused_prep_multi", arr, n, self.cur)

    def fused_prep(self, src: torch.Tensor, dst_lo: torch.Tensor, W0: int = 0, gs: int = 0, N: int = 0, K0: int = 0, G: int = 0,
                   w0pad: Optional[torch.Tensor] = None) -> Callable[[], None]:
        """dst_lo = src - trunc_tf32(src) over a whole parameter arena, and the zero-padded [2][G][N][32] copy (+ lo words)
        of the first layer's weights: the operand copies the fused passes fetch by TMA."""
        sp, dp, n = C.c_void_p(src.data_ptr()), C.c_void_p(dst_lo.data_ptr()), src.numel()
        wp = C.c_void_p(w0pad.data_ptr()) if w0pad is not None else None
        w0 = C.c_void_p(W0) if W0 else None
        keep = (src, dst_lo, w0pad)
        return lambda keep=keep: L.call("orlk_fused_prep", sp, dp, n, w0, gs, N, K0, G, wp, self.cur)

    @staticmethod
    def effective_splits(K: int, want: int, cfg: int) -> int:
        """Number of k-splits the descriptor builder will actually produce for (K, want)."""
        BK = L.CFG_TILES[cfg][2]
        want = max(1, want)
        chunk = -(-(-(-K // want)) // BK) * BK
        return max(1, -(-K // chunk))

    def concat(self, segs: Sequence[Tuple[Mat, Mat, int, Mat]]) -> Callable[[], None]:
        """segs: (dst, src1, rep1, src2): dst[m] = [src1[m // rep1] | src2[m]]."""
        arr = (L.ConcatSeg * len(segs))()
        row = 0
        for i, (dst, s1, rep, s2) in enumerate(segs):
            sg = arr[i]
            sg.dst, sg.src1, sg.src2 = dst.ptr, s1.ptr, s2.ptr
            sg.ld_dst, sg.ld1, sg.ld2 = dst.ld, s1.ld, s2.ld
            sg.M, sg.w1, sg.w2, sg.rep1, sg.row_start = dst.rows, s1.cols, s2.cols, rep, row
            assert dst.cols == s1.cols + s2.cols and s2.rows == dst.rows and s1.rows * rep >= dst.rows
            row += dst.rows
        dev = self.upload_bytes(bytes(arr))
        n, total, ptr = len(segs), row, C.c_void_p(dev.data_ptr())
        return lambda: L.call("orlk_concat_rows", ptr, n, total, self.cur)

    def adam(self, descs: Sequence["AdamT"], groups_ptr: int) -> Callable[[], None]:
        arr = (L.AdamDesc * len(descs))()
        blk = 0
        for i, t in enumerate(descs):
            d = arr[i]
            d.p, d.m, d.v, d.tgt, d.grad = t.p, t.m or None, t.v or None, t.tgt or None, t.grad or None
            d.n, d.g_split_stride, d.g_splits, d.group = t.n, t.g_split_stride, t.g_splits, t.group
            d.wd, d.block_start, d.flags = t.wd, blk, t.flags
            d.pT, d.cols = t.pT or None, max(t.cols, 1)
            blk += -(-t.n // 128)
        dev = self.upload_bytes(bytes(arr))
        n, total, ptr, gp = len(descs), blk, C.c_void_p(dev.data_ptr()), C.c_void_p(groups_ptr)
        return lambda: L.call("orlk_adam_step", ptr, n, total, gp, self.cur)


@dataclass
class AdamT:
    p: int
    n: int
    group: int
    m: int = 0
    v: int = 0
    tgt: int = 0
    grad: int = 0
    g_splits: int = 1
    g_split_stride: int = 0
    wd: float = 0.0
    flags: int = L.OPT_ADAM
    pT: int = 0
    cols: int = 1


class Plan:
    """An ordered list of launches; runs eagerly (debug) or as one captured CUDA graph.

    ``fork()`` / ``branch(k)`` / ``join()`` mark launches that are independent of each other: branch k > 0 runs on a
    side stream that waits for everything issued before the fork, and ``join`` makes the main stream wait for all
    branches.  Under stream capture this becomes parallel branches of the graph (the GPU runs e.g. the weight
    gradients of different layers concurrently instead of one partially filled launch after another)."""

    N_SIDE = 3

    def __init__(self, rt: Runtime, name: str = ""):
        self.rt, self.name = rt, name
        self.ops: List[Tuple[str, Callable[[], None]]] = []
        self.graph: Optional[C.c_void_p] = None
        # The launch closures hold raw device pointers: every tensor they address must be referenced from here (or
        # from the owning learner) for as long as the plan can run.
        self.keep: List[object] = []
        self._branch = 0
        self._side = None
        self._ev_fork = None
        self._ev_join = None
        self.flat_ops: List[Tuple[str, Callable[[], None]]] = []      # every launch, branch-agnostic (profiling)

    def add(self, label: str, op: Callable[[], None]) -> None:
        b = self._branch
        self.flat_ops.append((label, op))
        if b == 0:
            self.ops.append((label, op))
        else:
            self.ops.append((label, lambda op=op, b=b: self._on_side(b, op)))

    # ---- parallel sections
    def _ensure_side(self) -> None:
        if self._side is None:
            self._side, self._ev_join = [], []
            for _ in range(self.N_SIDE):
                st, ev = C.c_void_p(), C.c_void_p()
                L.call("orlk_stream_create", C.byref(st))
                L.call("orlk_event_create_notiming", C.byref(ev))
                self._side.append(st)
                self._ev_join.append(ev)
            self._ev_fork = C.c_void_p()
            L.call("orlk_event_create_notiming", C.byref(self._ev_fork))

    def _on_side(self, b: int, op: Callable[[], None]) -> None:
        rt = self.rt
        main = rt.cur
        rt.cur = self._side[b - 1]
        try:
            op()
        finally:
            rt.cur = main

    def fork(self) -> None:
        self._ensure_side()

        def op():
            L.call("orlk_event_record", self._ev_fork, self.rt.cur)
            for st in self._side:
                L.call("orlk_stream_wait_event", st, self._ev_fork)
        self.ops.append(("fork", op))
        self._branch = 0

    def branch(self, k: int) -> None:
        assert 0 <= k <= self.N_SIDE
        self._branch = k

    def join(self) -> None:
        def op():
            for st, ev in zip(self._side, self._ev_join):
                L.call("orlk_event_record", ev, st)
                L.call("orlk_stream_wait_event", self.rt.cur, ev)
        self.ops.append(("join", op))
        self._branch = 0

    # ---- point-to-point dependencies between branches (what fork / join cannot express)
    def record_event(self) -> C.c_void_p:
        """An event recorded at the current position of the current branch; ``wait_event`` makes another branch wait for it."""
        ev = C.c_void_p()
        L.call("orlk_event_create_notiming", C.byref(ev))
        self.keep.append(ev)
        b = self._branch
        op = lambda: L.call("orlk_event_record", ev, self.rt.cur)
        self.ops.append(("fork", (lambda: self._on_side(b, op)) if b else op))
        return ev

    def wait_event(self, ev: C.c_void_p) -> None:
        b = self._branch
        op = lambda: L.call("orlk_stream_wait_event", self.rt.cur, ev)
        self.ops.append(("join", (lambda: self._on_side(b, op)) if b else op))

    # ---- a detached launch: runs on its own stream after everything issued so far, joined by join_detached()
    @property
    def has_detached(self) -> bool:
        return getattr(self, "_aux", None) is not None

    def detach(self, label: str, op: Callable[[], None]) -> None:
        if not self.has_detached:
            self._aux, self._ev_aux_fork, self._ev_aux_join = C.c_void_p(), C.c_void_p(), C.c_void_p()
            L.call("orlk_stream_create", C.byref(self._aux))
            L.call("orlk_event_create_notiming", C.byref(self._ev_aux_fork))
            L.call("orlk_event_create_notiming", C.byref(self._ev_aux_join))

        def run():
            rt = self.rt
            L.call("orlk_event_record", self._ev_aux_fork, rt.cur)
            L.call("orlk_stream_wait_event", self._aux, self._ev_aux_fork)
            main = rt.cur
            rt.cur = self._aux
            try:
                op()
            finally:
                rt.cur = main
        self.flat_ops.append((label, op))
        self.ops.append((label, run))

    def join_detached(self) -> None:
        def op():
            L.call("orlk_event_record", self._ev_aux_join, self._aux)
            L.call("orlk_stream_wait_event", self.rt.cur, self._ev_aux_join)
        self.ops.append(("join", op))

    @property
    def n_launches(self) -> int:
        return sum(1 for lbl, _ in self.ops if lbl not in ("fork", "join"))

    def run_eager(self) -> None:
        for _, op in self.ops:
            op()

    def capture(self) -> None:
        g = C.c_void_p()
        rt = self.rt
        torch.cuda.synchronize(rt.device)
        # No CUDA call other than the captured launches may come from this thread while the capture is open: a finaliser that
        # the cyclic collector happens to run in the middle (an old Plan destroying its graph) invalidates the capture
        # silently.  So: pending destructions first, the collector off for the duration, late finalisers deferred.
        _flush_deferred_graphs()
        gc_was_on = gc.isenabled()
        gc.disable()
        _CAPTURING[0] = True
        rt.cur = C.c_void_p(rt.capture_stream.cuda_stream)
        try:
            L.call("orlk_graph_begin", rt.cur)
            try:
                if os.environ.get("ORLK_GRAPH_DEBUG", "0") == "2":      # which launch breaks a capture
                    cap = C.c_void_p(rt.cur.value)
                    for lbl, op in self.ops:
                        op()
                        st = L.load().orlk_capture_status(cap)
                        if st != 1:
                            raise L.OrlkError(f"capture of plan {self.name!r} left state {st} behind op {lbl!r}")
                else:
                    self.run_eager()
            finally:
                L.call("orlk_graph_end", rt.cur, C.byref(g))
        finally:
            rt.cur = rt.exec_ptr
            _CAPTURING[0] = False
            if gc_was_on:
                gc.enable()
        self.graph = g

    def launch(self) -> None:
        if self.graph is None:
            self.capture()
        L.call("orlk_graph_launch", self.graph, self.rt.cur)

    def __del__(self):
        try:
            if self.graph is not None:
                if _CAPTURING[0]:
                    _DEFERRED_GRAPHS.append(self.graph)
                else:
                    L.load().orlk_graph_destroy(self.graph)
        except Exception:
            pass


_CAPTURING = [False]
_DEFERRED_GRAPHS: List[C.c_void_p] = []


def _flush_deferred_graphs() -> None:
    while _DEFERRED_GRAPHS:
        try:
            L.load().orlk_graph_destroy(_DEFERRED_GRAPHS.pop())
        except Exception:
            pass


_RUNTIMES: Dict[str, Runtime] = {}


def get_runtime(device) -> Runtime:
    """One Runtime per CUDA device, shared by buffers, policies and dynamics."""
    dev = require_cuda(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    key = f"cuda:{idx}"
    if key not in _RUNTIMES:
        _RUNTIMES[key] = Runtime(torch.device(key))
    return _RUNTIMES[key]
