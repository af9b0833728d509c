"""Parameter arenas and launch emitters for dense (MLP / ensemble) networks.

A ``ParamSet`` owns the parameters of G identically-shaped networks ("members": twin critics, or the E members
of an EnsembleLinear stack), their Adam moments and (optionally) their target copies, each in ONE contiguous
fp32 device tensor.  The ``nn.Parameter`` objects of the facade modules are re-pointed at views of that tensor,
so ``state_dict()`` / ``torch.save`` / ``load_state_dict`` keep working unchanged while the
kernels see ``base + offset``.
"""
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _lib as L
from .core import GP, AdamT, Mat, Plan, Runtime, NUM_SMS


@dataclass
class Layer:
    """One dense layer of a member group.  ``layout`` is 'oi' (nn.Linear, W[out,in]) or 'io' (EnsembleLinear W[in,out])."""
    in_dim: int
    out_dim: int
    layout: str
    w_off: int          # offsets (floats) of member 0 inside the ParamSet block
    b_off: int
    w_gs: int           # member stride (floats)
    b_gs: int

    @property
    def w_numel(self) -> int:
        return self.in_dim * self.out_dim


class ParamSet:
    def __init__(self, rt: Runtime, name: str):
        self.rt, self.name = rt, name
        self.G = 0
        self.total = 0                    # floats in the parameter tensor
        self.layers: List[Layer] = []     # hidden layers followed by narrow heads, in forward order
        self.extra: Dict[str, Tuple[int, int]] = {}   # name -> (offset, numel) of non-layer parameters
        self.P = self.Mo = self.Vo = self.T = None
        self.WT = None                    # transposed weight copies (K-major operand of the tensor-core dgrad)
        self.wt_layers: List[int] = []
        self._wt_version = -1
        self._adopted: List[torch.Tensor] = []   # the facade parameters whose storage is this arena (see _host_version)
        self.group_ids: List[int] = []     # Adam group index per member (set by the learner)

    # ------------------------------------------------------------------ construction from facade modules
    @staticmethod
    def _align(n: int) -> int:
        return (n + 3) // 4 * 4

    @classmethod
    def from_linear_members(cls, rt: Runtime, name: str, members: Sequence[Sequence[nn.Linear]],
                            targets: Optional[Sequence[Sequence[nn.Linear]]] = None,
                            extra: Optional[Sequence[Dict[str, nn.Parameter]]] = None,
                            fuse_last: int = 1) -> "ParamSet":
        """members[g] = the nn.Linear modules of member g in forward order.  The last ``fuse_last`` Linears are
        narrow heads; when ``fuse_last`` == 2 (dist_net.mu, dist_net.sigma) they are placed back to back so the
        engine sees one [2A, K] head."""
        self = cls(rt, name)
        self.G = len(members)
        ref = members[0]
        off = 0
        n_lin = len(ref)
        spec = []
        i = 0
        while i < n_lin:
            lin = ref[i]
            if fuse_last == 2 and i == n_lin - 2:
                nxt = ref[i + 1]
                assert nxt.in_features == lin.in_features
                out = lin.out_features + nxt.out_features
                w_off = off; off += out * lin.in_features
                off = cls._align(off)
                b_off = off; off += out
                off = cls._align(off)
                spec.append((lin.in_features, out, w_off, b_off, (i, i + 1)))
                i += 2
            else:
                w_off = off; off += lin.out_features * lin.in_features
                off = cls._align(off)
                b_off = off; off += lin.out_features
                off = cls._align(off)
                spec.append((lin.in_features, lin.out_features, w_off, b_off, (i,)))
                i += 1
        extra_spec = {}
        if extra is not None:
            for k, p in extra[0].items():
                extra_spec[k] = (off, p.numel())
                off = cls._align(off + p.numel())
        block = cls._align(off)
        self.total = block * self.G
        self.block = block
        self.layers = [Layer(i_, o_, "oi", w, b, block, block) for (i_, o_, w, b, _) in spec]
        self.extra = extra_spec
        self._alloc(targets is not None)
        for g, lins in enumerate(members):
            self._adopt_linears(self.P, g, lins, spec)
            if extra is not None:
                for k, p in extra[g].items():
                    o, n = extra_spec[k]
                    self._adopt(self.P, g * block + o, p)
        if targets is not None:
            for g, lins in enumerate(targets):
                self._adopt_linears(self.T, g, lins, spec)
        return self

    def _adopt_linears(self, store, g, lins, spec):
        for (in_f, out_f, w_off, b_off, idxs) in spec:
            wo, bo = g * self.block + w_off, g * self.block + b_off
            for j in idxs:
                lin = lins[j]
                self._adopt(store, wo, lin.weight)
                self._adopt(store, bo, lin.bias)
                wo += lin.weight.numel()
                bo += lin.bias.numel()

    @classmethod
    def from_ensemble(cls, rt: Runtime, name: str, layers: Sequence[nn.Module],
                      targets: Optional[Sequence[nn.Module]] = None,
                      extra: Optional[Dict[str, nn.Parameter]] = None) -> "ParamSet":
        """layers = EnsembleLinear-like modules (``weight`` [E,in,out], ``bias`` [E,1,out]) in forward order."""
        self = cls(rt, name)
        E = layers[0].weight.shape[0]
        self.G = E
        off = 0
        spec = []
        for lay in layers:
            _, i_, o_ = lay.weight.shape
            w_off = off; off = cls._align(off + E * i_ * o_)
            b_off = off; off = cls._align(off + E * o_)
            spec.append((i_, o_, w_off, b_off))
        extra_spec = {}
        if extra is not None:
            for k, p in extra.items():
                extra_spec[k] = (off, p.numel())
                off = cls._align(off + p.numel())
        self.total = self.block = off
        self.layers = [Layer(i_, o_, "io", w, b, i_ * o_, o_) for (i_, o_, w, b) in spec]
        self.extra = extra_spec
        self._alloc(targets is not None)
        for lay, (i_, o_, w, b) in zip(layers, spec):
            self._adopt(self.P, w, lay.weight)
            self._adopt(self.P, b, lay.bias)
        if extra is not None:
            for k, p in extra.items():
                self._adopt(self.P, extra_spec[k][0], p)
        if targets is not None:
            for lay, (i_, o_, w, b) in zip(targets, spec):
                self._adopt(self.T, w, lay.weight)
                self._adopt(self.T, b, lay.bias)
        return self

    def _alloc(self, with_target: bool) -> None:
        self.P = self.rt.zeros(self.total)
        self.Mo = self.rt.zeros(self.total)
        self.Vo = self.rt.zeros(self.total)
        self.T = self.rt.zeros(self.total) if with_target else None

    def _adopt(self, store: torch.Tensor, off: int, p: torch.Tensor) -> None:
        """Copy the parameter's current value into the arena and re-point its storage at the arena view."""
        view = store[off:off + p.numel()].view(p.shape)
        with torch.no_grad():
            view.copy_(p.detach().to(store.device, torch.float32))
        p.data = view
        if store is self.P:
            self._adopted.append(p)

    def _host_version(self) -> int:
        """Changes whenever torch code writes a parameter in place (load_state_dict's ``param.copy_``, ``nn.init``,
        ``p.mul_``).  ``p.data = view`` gives every adopted parameter its OWN version counter, so the arena tensor's
        counter alone does not see those writes; the engine's kernels write through raw pointers and keep the
        derived copies in sync themselves.  Writes through ``p.data`` bump nothing: call ``invalidate()`` after those."""
        return self.P._version + sum(p._version for p in self._adopted)

    def invalidate(self) -> None:
        """Force the next ``refresh_wt`` to re-derive the transposed weight copies from the parameters."""
        self._wt_version = -1

    # ------------------------------------------------------------------ transposed weight copies
    def enable_wt(self, layers: Sequence[int]) -> None:
        """Maintain WT[l] = W[l]^T ([in, out] row-major) for the given 'oi' layers.  The Adam kernel keeps them in
        sync after every update (OrlkAdamDesc.pT); ``refresh_wt`` re-derives them after host-side writes."""
        new = [l for l in layers if l not in self.wt_layers]
        if not new:
            return
        if self.WT is None:
            self.WT = self.rt.zeros(self.total)
        self.wt_layers += new
        self._wt_version = -1
        self.refresh_wt()

    def refresh_wt(self) -> None:
        if self.WT is None or self._host_version() == self._wt_version:
            return
        with torch.no_grad():
            for l in self.wt_layers:
                lay = self.layers[l]
                assert lay.layout == "oi"
                for g in range(self.G):
                    o = lay.w_off + g * lay.w_gs
                    src = self.P[o:o + lay.w_numel].view(lay.out_dim, lay.in_dim)
                    self.WT[o:o + lay.w_numel].view(lay.in_dim, lay.out_dim).copy_(src.t())
        self._wt_version = self._host_version()

    # ------------------------------------------------------------------ lo operand words (fused tensor-core passes)
    def lo_arena(self, store: str) -> torch.Tensor:
        """Arena shaped like ``store`` ("P", "T" or "WT") holding x - trunc_tf32(x); refreshed inside the step graph by
        an ``orlk_fused_prep`` launch (emit_lo_refresh), never on the host."""
        name = store + "lo"
        t = getattr(self, name, None)
        if t is None:
            t = self.rt.zeros(self.total)
            setattr(self, name, t)
        return t

    def w_lo(self, l: int, g: int = 0, store: str = "P") -> int:
        lay = self.layers[l]
        return self._ptr(self.lo_arena(store), lay.w_off + g * lay.w_gs)

    def w0_pad(self, store: str) -> torch.Tensor:
        """[2][G][out][32]: the first layer's [out][in <= 32] weights zero-padded to 32 columns (TMA cannot address rows of
        ``in`` floats) and their lo words; refreshed together with ``lo_arena`` (orlk_fused_prep)."""
        name = store + "w0pad"
        t = getattr(self, name, None)
        if t is None:
            t = self.rt.zeros(2, self.G, self.layers[0].out_dim, 32)
            setattr(self, name, t)
        return t

    def wt(self, l: int, g: int = 0) -> int:
        lay = self.layers[l]
        return self._ptr(self.WT, lay.w_off + g * lay.w_gs)

    # ------------------------------------------------------------------ pointers
    def _ptr(self, store: torch.Tensor, off: int) -> int:
        return store.data_ptr() + 4 * off

    def w(self, l: int, g: int = 0, store: str = "P") -> int:
        lay = self.layers[l]
        return self._ptr(getattr(self, store), lay.w_off + g * lay.w_gs)

    def b(self, l: int, g: int = 0, store: str = "P") -> int:
        lay = self.layers[l]
        return self._ptr(getattr(self, store), lay.b_off + g * lay.b_gs)

    def extra_ptr(self, name: str, store: str = "P") -> int:
        return self._ptr(getattr(self, store), self.extra[name][0])


class GradBuf:
    """Split-K partial gradient storage for one ParamSet: [n_slots, total] floats, zero-initialised.

    Slot s of a tensor lives at ``base + s*total + offset`` so that a single Adam descriptor per tensor can sum
    ``g_splits`` partials with stride ``total``."""

    def __init__(self, rt: Runtime, ps: ParamSet, n_slots: int):
        self.ps, self.n_slots = ps, n_slots
        self.buf = rt.zeros(n_slots * ps.total)

    def ptr(self, off: int) -> int:
        return self.buf.data_ptr() + 4 * off

    @property
    def stride(self) -> int:
        return self.ps.total


# ---------------------------------------------------------------------------------------------- emitters
# below this many rows a pass runs on the small-row kernel: a 1024-row layer is only 8 (x members) 128-row tensor-core
# tiles - 17-20 us on a handful of SMs against ~9 us as 512 small tiles (IQL at batch 1024)
TC_MIN_ROWS = int(os.environ.get("ORLK_TC_MIN_ROWS", "2048"))
# forward / dgrad: below this the layer is latency-bound and goes to the small-row fp32 kernel (csrc/orlk_tiny.cu);
# ORLK_TC_MIN_ROWS_FWD=128 restores the n-tiled tensor-core path for those layers
TC_MIN_ROWS_FWD = int(os.environ.get("ORLK_TC_MIN_ROWS_FWD", str(TC_MIN_ROWS)))


def tc_n_tile(M: int, N: int) -> int:
    """Output columns per CTA: everything for large M (one 128 x N accumulator), 32-column strips for small M."""
    return 0 if M >= TC_MIN_ROWS or N % 32 != 0 else 32


# Ensembles (EnsembleLinear, 'io' weights [member][in][out]) of short members: the small-row kernel costs ~2.5 us per
# member and layer (ten 256-row critics: 25 us), an n-tiled tensor-core launch ~11 us for all members together.
ENSEMBLE_TC = os.environ.get("ORLK_ENSEMBLE_TC", "1") != "0"


def ens_tc_ok(lays: Sequence["Layer"], G: int, M: int) -> bool:
    return ENSEMBLE_TC and G >= 4 and G * M >= TC_MIN_ROWS // 2 and M >= 128 and all(lay.layout == "io" for lay in lays)


def ens_n_tile(G: int, M: int, N: int) -> int:
    """Output columns per CTA for an ensemble launch: the narrowest 32-multiple that keeps the grid within one wave of
    148 CTAs; ensembles too large for that (EDAC's 50 critics on hopper) take the widest tile, i.e. the fewest waves."""
    best = 0
    for nt in (32, 64, 128, 256):
        if N % nt == 0:
            best = nt
            if G * (-(-M // 128)) * (N // nt) <= 148:
                return nt
    return best


def tc_ok_fwd_io(lay: Layer) -> bool:
    return lay.layout == "io" and lay.in_dim % 4 == 0 and lay.out_dim % 32 == 0 and lay.out_dim <= 256 and lay.w_gs % 4 == 0


def tc_ok_dgrad_io(lay: Layer) -> bool:
    return lay.layout == "io" and lay.out_dim % 4 == 0 and lay.in_dim % 32 == 0 and lay.in_dim <= 256 and lay.w_gs % 4 == 0


def tc_ok_fwd(lay: Layer, M: int) -> bool:
    return lay.layout == "oi" and lay.in_dim % 4 == 0 and lay.out_dim % 16 == 0 and lay.out_dim <= 256 and M >= TC_MIN_ROWS_FWD


def tc_ok_dgrad(lay: Layer, M: int) -> bool:
    return lay.layout == "oi" and lay.out_dim % 4 == 0 and lay.in_dim % 16 == 0 and lay.in_dim <= 256 and M >= TC_MIN_ROWS_FWD


def tc_ok_wgrad(lay: Layer, M: int) -> bool:
    return lay.layout == "oi" and lay.in_dim % 16 == 0 and lay.in_dim <= 256 and lay.out_dim >= 64 and M >= TC_MIN_ROWS


def pick_cfg(M: int, N: int, rows_per_problem: Optional[int] = None) -> int:
    """Tile configuration for a launch whose problems add up to M x N outputs (see csrc/orlk_gemm.cu).
    ``rows_per_problem``: when every problem of the launch is short (an ensemble of 256-row members), the small-row
    kernel wins however many members there are: measured 38 us -> 15 us for ten 256 x 256 x 256 members."""
    if rows_per_problem is not None and rows_per_problem < TC_MIN_ROWS and os.environ.get("ORLK_ENSEMBLE_TINY", "1") != "0":
        return L.CFG_TINY
    if M * N >= 128 * 128 * 96:
        return L.CFG_BIG
    if M * N >= 64 * 64 * 96:
        return L.CFG_MID
    return L.CFG_TINY      # small layers are latency-bound: 32 x 16 tiles, whole-k cp.async burst, 4 k-parallel groups


def fwd_problem(ps: ParamSet, l: int, g: int, X: Mat, Y: Mat, epi: int, store: str = "P", Z: Optional[Mat] = None,
                YT: Optional[Mat] = None) -> GP:
    """Y = act(X W^T + b) for member g of layer l (nn.Linear, mlp.py:22 / EnsembleLinear, ensemble_linear.py:30-41)."""
    lay = ps.layers[l]
    assert X.cols == lay.in_dim and Y.cols == lay.out_dim and X.rows == Y.rows, (X, Y, lay)
    if lay.layout == "oi":
        b_layout, ldb = 1, lay.in_dim
    else:
        b_layout, ldb = 0, lay.out_dim
    return GP(A=X.ptr, lda=X.ld, a_layout=0, B=ps.w(l, g, store), ldb=ldb, b_layout=b_layout, C=Y.ptr, ldc=Y.ld,
              M=X.rows, N=lay.out_dim, K=lay.in_dim, epi=epi, bias=ps.b(l, g, store),
              C2=Z.ptr if Z is not None else 0, CT=YT.ptr if YT is not None else 0, ldct=YT.ld if YT is not None else 0)


def dgrad_problem(ps: ParamSet, l: int, g: int, dY: Mat, dX: Mat, epi: int, aux: Optional[Mat],
                  col0: int = 0, ncols: Optional[int] = None, dXT: Optional[Mat] = None) -> GP:
    """dX[:, col0:col0+ncols] = (dY W)[:, col0:...] (x) mask  -- input gradient of layer l for member g."""
    lay = ps.layers[l]
    ncols = lay.in_dim - col0 if ncols is None else ncols
    assert dY.cols == lay.out_dim and dX.cols == ncols and dX.rows == dY.rows
    if lay.layout == "oi":      # W[out,in]: B(k=o, n=i) = W[o*in + i]
        Bp, b_layout, ldb = ps.w(l, g) + 4 * col0, 0, lay.in_dim
    else:                       # W[in,out]: B(k=o, n=i) = W[i*out + o]
        Bp, b_layout, ldb = ps.w(l, g) + 4 * col0 * lay.out_dim, 1, lay.out_dim
    return GP(A=dY.ptr, lda=dY.ld, a_layout=0, B=Bp, ldb=ldb, b_layout=b_layout, C=dX.ptr, ldc=dX.ld,
              M=dY.rows, N=ncols, K=lay.out_dim, epi=epi, aux=aux.ptr if aux is not None else 0,
              ldaux=aux.ld if aux is not None else 0, CT=dXT.ptr if dXT is not None else 0,
              ldct=dXT.ld if dXT is not None else 0)


def wgrad_problem(ps: ParamSet, gb: GradBuf, l: int, g: int, X: Mat, dY: Mat, k_splits: int, split_base: int = 0,
                  col0: int = 0, with_bias: bool = True) -> GP:
    """Partial dW (and db) of layer l, member g, reduced over the rows of X / dY, into GradBuf slots."""
    lay = ps.layers[l]
    assert X.rows == dY.rows and dY.cols == lay.out_dim
    w_off = lay.w_off + g * lay.w_gs
    b_off = lay.b_off + g * lay.b_gs
    if lay.layout == "oi":      # dW[o,i] = sum_m dY[m,o] X[m,i]
        return GP(A=dY.ptr, lda=dY.ld, a_layout=1, B=X.ptr, ldb=X.ld, b_layout=0, C=gb.ptr(w_off + col0), ldc=lay.in_dim,
                  M=lay.out_dim, N=X.cols, K=X.rows, k_splits=k_splits, split_base=split_base,
                  c_split_stride=gb.stride, sum_split_stride=gb.stride,
                  rowsum=gb.ptr(b_off) if with_bias else 0)
    return GP(A=X.ptr, lda=X.ld, a_layout=1, B=dY.ptr, ldb=dY.ld, b_layout=0, C=gb.ptr(w_off + col0 * lay.out_dim),
              ldc=lay.out_dim, M=X.cols, N=lay.out_dim, K=X.rows, k_splits=k_splits, split_base=split_base,
              c_split_stride=gb.stride, sum_split_stride=gb.stride, colsum=gb.ptr(b_off) if with_bias else 0)


def adam_descs(ps: ParamSet, gb: GradBuf, splits_per_layer: Sequence[int], polyak: bool,
               layers: Optional[Sequence[int]] = None, grad_src: Optional[Dict] = None,
               members: Optional[Sequence[int]] = None) -> List[AdamT]:
    """One Adam(+polyak) descriptor per weight / bias tensor (all members of a tensor when they are contiguous).

    ``grad_src[(layer, 'w'|'b')] = (ptr of member 0, member stride, split stride, n splits)`` overrides the GradBuf
    location for tensors whose partial gradients were written elsewhere (the narrow-layer kernels)."""
    out = []
    grad_src = grad_src or {}
    idx = range(len(ps.layers)) if layers is None else layers
    for l in idx:
        lay, s = ps.layers[l], splits_per_layer[l]
        contiguous = (lay.w_gs == lay.w_numel)          # ensemble tensors: members back to back
        for g in ([0] if contiguous else (members if members is not None else range(ps.G))):
            nw = lay.w_numel * (ps.G if contiguous else 1)
            nb = lay.out_dim * (ps.G if contiguous else 1)
            wo = lay.w_off + (0 if contiguous else g * lay.w_gs)
            bo = lay.b_off + (0 if contiguous else g * lay.b_gs)
            for off, n, is_w in ((wo, nw, True), (bo, nb, False)):
                keep_t = is_w and ps.WT is not None and l in ps.wt_layers and not contiguous
                src = grad_src.get((l, "w" if is_w else "b"))
                if src is not None:
                    gptr, gsplits, gstride = src[0] + 4 * g * src[1], src[3], src[2]
                else:
                    gptr, gsplits, gstride = gb.ptr(off), s, gb.stride
                out.append(AdamT(p=ps._ptr(ps.P, off), n=n, group=ps.group_ids[g], m=ps._ptr(ps.Mo, off),
                                 v=ps._ptr(ps.Vo, off), tgt=ps._ptr(ps.T, off) if (polyak and ps.T is not None) else 0,
                                 grad=gptr, g_splits=gsplits, g_split_stride=gstride,
                                 flags=L.OPT_ADAM | (L.OPT_POLYAK if polyak and ps.T is not None else 0),
                                 pT=ps._ptr(ps.WT, off) if keep_t else 0, cols=lay.in_dim if keep_t else 1))
    return out


def wgrad_splits(out_tiles: int, K: int, cfg: int, target_ctas: Optional[int] = None) -> int:
    """Split-K factor so that a wgrad launch fills the machine: ~target_ctas CTAs, chunks >= 2*BK."""
    BK = L.CFG_TILES[cfg][2]
    if target_ctas is None:
        target_ctas = NUM_SMS if cfg == L.CFG_BIG else 2 * NUM_SMS
    want = max(1, target_ctas // max(1, out_tiles))
    want = min(want, max(1, K // (2 * BK)))
    return Runtime.effective_splits(K, want, cfg)
