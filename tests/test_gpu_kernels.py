"""GPU: every kernel entry point of the C ABI against a plain PyTorch fp32/fp64 expression of the same op."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def rt():
    from offlinerlkit_b200.engine.core import get_runtime
    return get_runtime(DEV)


def _close(got, ref, rtol=1e-5, atol=1e-5, msg=""):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    scale = ref.abs().max().item() if ref.numel() else 1.0
    assert err <= atol + rtol * scale, f"{msg}: max err {err:.3e} (scale {scale:.3e})"


# ------------------------------------------------------------------------------------------------ GEMM
def _gemm_case(rt, cfg, M, N, K, a_layout, b_layout, epi, splits, with_bias, sums, seed, passes=0):
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import GP
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=gen)
    B = torch.randn(K, N, generator=gen)
    bias = torch.randn(N, generator=gen) if with_bias else None
    aux = torch.randn(M, N, generator=gen)
    Ad = (A if a_layout == 0 else A.t().contiguous()).to(DEV)
    Bd = (B if b_layout == 0 else B.t().contiguous()).to(DEV)
    s_eff = rt.effective_splits(K, splits, cfg)
    Cd = torch.full((s_eff, M, N), float("nan"), device=DEV)
    C2 = torch.zeros(M, N, device=DEV)
    W = max(M, N)
    rs = torch.full((s_eff, W), float("nan"), device=DEV)
    cs = torch.full((s_eff, W), float("nan"), device=DEV)
    p = GP(A=Ad.data_ptr(), lda=Ad.stride(0), a_layout=a_layout, B=Bd.data_ptr(), ldb=Bd.stride(0), b_layout=b_layout,
           C=Cd.data_ptr(), ldc=N, M=M, N=N, K=K, epi=epi, bias=bias.to(DEV).data_ptr() if with_bias else 0,
           aux=0, ldaux=N, C2=C2.data_ptr() if epi == L.EPI_SWISH else 0, k_splits=splits, c_split_stride=M * N,
           sum_split_stride=0)
    keep = [bias.to(DEV) if with_bias else None]
    if with_bias:
        p.bias = keep[0].data_ptr()
    auxd = aux.to(DEV)
    if epi in (L.EPI_RELU_MASK, L.EPI_DSWISH):
        p.aux = auxd.data_ptr()
    if sums:
        p.rowsum, p.colsum, p.sum_split_stride = rs.data_ptr(), cs.data_ptr(), W
    rt.gemm([p], cfg, passes=passes)()
    torch.cuda.synchronize()
    ref = A.double() @ B.double()
    if with_bias:
        ref = ref + bias.double()
    got = Cd.sum(0) if epi == L.EPI_NONE else Cd[0]
    if epi == L.EPI_RELU:
        ref = ref.clamp(min=0)
    elif epi == L.EPI_RELU_MASK:
        ref = ref * (aux > 0)
    elif epi == L.EPI_SWISH:
        _close(C2, ref, msg="swish z", **(dict(rtol=2e-3, atol=2e-3 * math.sqrt(K)) if passes == 1 else {}))
        ref = ref * torch.sigmoid(ref)
    elif epi == L.EPI_DSWISH:
        s = torch.sigmoid(aux.double())
        ref = ref * (s * (1 + aux.double() * (1 - s)))
    tag = f"cfg{cfg} {M}x{N}x{K} a{a_layout} b{b_layout} epi{epi} s{splits} passes{passes}"
    assert not torch.isnan(got).any(), tag + ": NaN = unwritten output"
    if passes == 1:     # single-pass TF32: 10-bit mantissas
        _close(got, ref, rtol=2e-3, atol=2e-3 * math.sqrt(K), msg=tag)
    else:
        _close(got, ref, rtol=2e-6 * math.sqrt(K) + 1e-6, atol=1e-5, msg=tag)
    if sums:
        _close(rs[:, :M].sum(0), A.double().sum(1), rtol=1e-5, atol=1e-4, msg=tag + " rowsum")
        _close(cs[:, :N].sum(0), B.double().sum(0), rtol=1e-5, atol=1e-4, msg=tag + " colsum")


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
def test_gemm_layouts_and_ragged_shapes(rt, cfg):
    shapes = [(1, 1, 1), (37, 23, 23), (300, 256, 250), (128, 128, 64), (129, 6, 256), (256, 256, 256), (64, 1, 515)]
    seed = 0
    for (M, N, K) in shapes:
        for a_layout in (0, 1):
            for b_layout in (0, 1):
                seed += 1
                _gemm_case(rt, cfg, M, N, K, a_layout, b_layout, 0, 1, with_bias=bool(seed % 2), sums=False, seed=seed)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
def test_gemm_epilogues(rt, cfg):
    for epi in (1, 2, 3, 4):
        _gemm_case(rt, cfg, 200, 136, 96, 0, 1, epi, 1, with_bias=epi in (1, 3), sums=False, seed=100 + epi)
        _gemm_case(rt, cfg, 256, 256, 256, 0, 0, epi, 1, with_bias=epi in (1, 3), sums=False, seed=200 + epi)


@pytest.mark.parametrize("cfg", [0, 2, 3])
def test_gemm_split_k_and_sums(rt, cfg):
    # wgrad shapes: C[out,in] = dY^T X reduced over rows, with bias row/col sums
    for (M, N, K, splits) in [(256, 256, 7936, 18), (256, 23, 7936, 18), (1, 256, 7936, 9), (12, 256, 256, 4),
                              (256, 17, 256, 3), (100, 50, 1000, 7)]:
        _gemm_case(rt, cfg, M, N, K, 1, 0, 0, splits, with_bias=False, sums=True, seed=M + N + K)


@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_small_row_kernel_tensor_core_variant(rt, passes):
    # mma.sync TF32 micro-kernel of the small-row GEMM: every layout, ragged shapes, epilogues, several k passes
    seed = 500
    for (M, N, K) in [(1, 1, 1), (37, 23, 23), (300, 256, 250), (129, 6, 256), (256, 256, 256), (64, 1, 515), (256, 256, 17)]:
        for a_layout in (0, 1):
            for b_layout in (0, 1):
                seed += 1
                _gemm_case(rt, 4, M, N, K, a_layout, b_layout, 0, 1, with_bias=bool(seed % 2), sums=False, seed=seed,
                           passes=passes)
    for epi in (1, 2, 3, 4):
        _gemm_case(rt, 4, 200, 136, 96, 0, 1, epi, 1, with_bias=epi in (1, 3), sums=False, seed=600 + epi, passes=passes)
        _gemm_case(rt, 4, 256, 256, 256, 0, 0, epi, 1, with_bias=epi in (1, 3), sums=False, seed=700 + epi, passes=passes)


@pytest.mark.parametrize("passes", [3, 1])
def test_gemm_chain_forward_and_backward(rt, passes):
    """The fused chain launch against the same layers in fp64: a 23 -> 256 -> 256 -> 256 -> 12 forward (bias + ReLU, no
    activation on the head) and the matching input-gradient chain (ReLU masks), two independent chains, ragged M."""
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import GP
    gen = torch.Generator().manual_seed(31)
    G, M, dims = 2, 200, [23, 256, 256, 256, 12]
    X = torch.randn(G, M, dims[0], generator=gen)
    Ws = [torch.randn(G, dims[i + 1], dims[i], generator=gen) / math.sqrt(dims[i]) for i in range(4)]
    bs = [torch.randn(G, dims[i + 1], generator=gen) * 0.1 for i in range(4)]
    Xd, Wd, bd = X.to(DEV), [w.to(DEV) for w in Ws], [b.to(DEV) for b in bs]
    H = [torch.full((G, M, dims[i + 1]), float("nan"), device=DEV) for i in range(4)]
    chains = []
    for g in range(G):
        st = []
        for i in range(4):
            src = Xd[g] if i == 0 else H[i - 1][g]
            st.append(GP(A=src.data_ptr(), lda=dims[i], a_layout=0, B=Wd[i][g].data_ptr(), ldb=dims[i], b_layout=1,
                         C=H[i][g].data_ptr(), ldc=dims[i + 1], M=M, N=dims[i + 1], K=dims[i], epi=L.EPI_RELU if i < 3 else L.EPI_NONE,
                         bias=bd[i][g].data_ptr()))
        chains.append(st)
    rt.gemm_chain(chains, passes)()
    torch.cuda.synchronize()
    tol = dict(rtol=3e-6, atol=3e-5) if passes == 3 else dict(rtol=5e-3, atol=5e-2)
    ref = X.double()
    refs = []
    for i in range(4):
        ref = torch.einsum("gmk,gnk->gmn", ref, Ws[i].double()) + bs[i].double()[:, None, :]
        if i < 3:
            ref = ref.clamp(min=0)
        refs.append(ref)
        assert not torch.isnan(H[i]).any(), f"chain fwd stage {i}: unwritten output"
        _close(H[i], ref, msg=f"chain fwd stage {i} passes {passes}", **tol)
    # backward: dOut [M,12] -> dZ2 = (dOut W3) * (H2 > 0) -> dZ1 -> dZ0, then d/d(input columns 17..22) without a mask
    dOut = torch.randn(G, M, 12, generator=gen)
    dOd = dOut.to(DEV)
    Hm = [r.float().to(DEV) for r in refs[:3]]          # exact masks from the reference activations
    dZ = [torch.full((G, M, 256), float("nan"), device=DEV) for _ in range(3)]
    dA = torch.full((G, M, 6), float("nan"), device=DEV)
    chains = []
    for g in range(G):
        st, src, kdim = [], dOd[g], 12
        for i in (3, 2, 1):
            st.append(GP(A=src.data_ptr(), lda=kdim, a_layout=0, B=Wd[i][g].data_ptr(), ldb=dims[i], b_layout=0,
                         C=dZ[i - 1][g].data_ptr(), ldc=256, M=M, N=256, K=kdim, epi=L.EPI_RELU_MASK, aux=Hm[i - 1][g].data_ptr(),
                         ldaux=256))
            src, kdim = dZ[i - 1][g], 256
        st.append(GP(A=dZ[0][g].data_ptr(), lda=256, a_layout=0, B=Wd[0][g].data_ptr() + 4 * 17, ldb=23, b_layout=0,
                     C=dA[g].data_ptr(), ldc=6, M=M, N=6, K=256, epi=L.EPI_NONE))
        chains.append(st)
    rt.gemm_chain(chains, passes)()
    torch.cuda.synchronize()
    gref = dOut.double()
    for i in (3, 2, 1):
        gref = torch.einsum("gmo,goi->gmi", gref, Ws[i].double()) * (refs[i - 1] > 0)
        assert not torch.isnan(dZ[i - 1]).any(), f"chain bwd stage {i}: unwritten output"
        _close(dZ[i - 1], gref, msg=f"chain bwd dZ{i - 1} passes {passes}", **tol)
    gA = torch.einsum("gmo,goi->gmi", gref, Ws[0].double()[:, :, 17:23])
    _close(dA, gA, msg=f"chain bwd d/da passes {passes}", **tol)


def test_gemm_small_row_kernel_sums_and_long_k(rt):
    # the small-row kernel (cfg 4) never splits k: whole reductions, several 256-wide passes, bias row/col sums
    for (M, N, K) in [(256, 256, 256), (256, 23, 256), (12, 256, 256), (256, 17, 700), (100, 50, 1000), (1, 256, 515)]:
        _gemm_case(rt, 4, M, N, K, 1, 0, 0, 1, with_bias=False, sums=True, seed=M + N + K)
        _gemm_case(rt, 4, M, N, K, 0, 1, 0, 1, with_bias=True, sums=True, seed=M + N + K + 1)


def test_gemm_grouped_many_problems(rt):
    from offlinerlkit_b200.engine.core import GP
    gen = torch.Generator().manual_seed(5)
    probs, refs, outs, keep = [], [], [], []
    for i, (M, N, K) in enumerate([(256, 256, 23), (40, 256, 256), (256, 1, 256), (77, 33, 129)] * 5):
        A, B = torch.randn(M, K, generator=gen).to(DEV), torch.randn(N, K, generator=gen).to(DEV)
        Cd = torch.zeros(M, N, device=DEV)
        probs.append(GP(A=A.data_ptr(), lda=K, a_layout=0, B=B.data_ptr(), ldb=K, b_layout=1, C=Cd.data_ptr(), ldc=N,
                        M=M, N=N, K=K))
        refs.append(A.double().cpu() @ B.double().cpu().t())
        outs.append(Cd)
        keep += [A, B]
    for cfg in (0, 1, 2, 3, 4):     # 20 problems: cfg 4 travels in kernel parameters, 16 problems per launch
        for o in outs:
            o.zero_()
        rt.gemm(probs, cfg)()
        torch.cuda.synchronize()
        for i, (o, r) in enumerate(zip(outs, refs)):
            _close(o, r, rtol=1e-5, atol=1e-4, msg=f"grouped cfg{cfg} problem {i}")


# ------------------------------------------------------------------------------------------------ narrow layers
def test_skinny_fwd_and_dgrad(rt):
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(1)
    for (G, M, K, NS) in [(1, 256, 256, 12), (2, 300, 256, 1), (3, 7, 40, 16), (2, 2560, 200, 6)]:
        X = torch.randn(G, M, K, generator=gen)
        W = torch.randn(G, NS, K, generator=gen)
        b = torch.randn(G, NS, generator=gen)
        Xd, Wd, bd = X.to(DEV), W.to(DEV), b.to(DEV)
        Y = torch.zeros(G, M, NS, device=DEV)
        L.call("orlk_skinny_fwd", Xd.data_ptr(), K, M * K, Wd.data_ptr(), K, 1, NS * K, bd.data_ptr(), NS, Y.data_ptr(), NS,
               M * NS, M, K, NS, G, rt.cur)
        # strided weight access (a column block of a [K, ld] matrix): Y = X @ Wk[:, c0:c0+NS]
        Wk = torch.randn(G, K, NS + 5, generator=gen)
        Wkd = Wk.to(DEV)
        Y2 = torch.zeros(G, M, NS, device=DEV)
        L.call("orlk_skinny_fwd", Xd.data_ptr(), K, M * K, Wkd.data_ptr() + 4 * 3, 1, NS + 5, K * (NS + 5), None, 0,
               Y2.data_ptr(), NS, M * NS, M, K, NS, G, rt.cur)
        _close(Y2, torch.einsum("gmk,gkn->gmn", X.double(), Wk.double()[:, :, 3:3 + NS]), msg=f"skinny fwd strided {G,M,K,NS}")
        ref = torch.einsum("gmk,gnk->gmn", X.double(), W.double()) + b.double()[:, None, :]
        _close(Y, ref, msg=f"skinny fwd {G,M,K,NS}")
        dY = torch.randn(G, M, NS, generator=gen)
        mask = torch.randn(G, M, K, generator=gen)
        dX = torch.zeros(G, M, K, device=DEV)
        dYd, md = dY.to(DEV), mask.to(DEV)
        dXT = torch.zeros(G, K, M, device=DEV)
        L.call("orlk_skinny_dgrad", dYd.data_ptr(), NS, M * NS, Wd.data_ptr(), K, NS * K, md.data_ptr(), K, M * K,
               dX.data_ptr(), K, M * K, dXT.data_ptr(), M, K * M, M, K, NS, G, rt.cur)
        ref = torch.einsum("gmn,gnk->gmk", dY.double(), W.double()) * (mask > 0)
        _close(dX, ref, msg=f"skinny dgrad {G,M,K,NS}")
        assert torch.equal(dXT, dX.transpose(1, 2).contiguous()), "transposed copy"
        dX2 = torch.zeros(G, M, K, device=DEV)
        L.call("orlk_skinny_dgrad", dYd.data_ptr(), NS, M * NS, Wd.data_ptr(), K, NS * K, None, 0, 0,
               dX2.data_ptr(), K, M * K, None, 0, 0, M, K, NS, G, rt.cur)
        _close(dX2, torch.einsum("gmn,gnk->gmk", dY.double(), W.double()), msg="skinny dgrad no mask")


def test_concat_rows(rt):
    from offlinerlkit_b200.engine.core import Mat
    obs, act = torch.randn(16, 5, device=DEV), torch.randn(64, 3, device=DEV)
    act0 = torch.randn(16, 3, device=DEV)
    X = torch.zeros(80, 8, device=DEV)
    Xm = Mat.of(X)
    rt.concat([(Xm.rows_(0, 16), Mat.of(obs), 1, Mat.of(act0)), (Xm.rows_(16, 80), Mat.of(obs), 4, Mat.of(act))])()
    torch.cuda.synchronize()
    ref = torch.cat([torch.cat([obs, act0], 1), torch.cat([obs.repeat_interleave(4, 0), act], 1)], 0)
    assert torch.equal(X, ref)


# ------------------------------------------------------------------------------------------------ policy head
def _head_ref(head, eps, A):
    mu, raw = head[:, :A], head[:, A:]
    sigma = raw.clamp(-5, 2).exp()
    u = mu + sigma * eps if eps is not None else mu
    a = torch.tanh(u)
    lp = (-((u - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - 0.5 * math.log(2 * math.pi)).sum(-1, keepdim=True)
    lp = lp - torch.log((1 - a.pow(2)) + 1e-6).sum(-1, keepdim=True)
    return a, lp


def _lp_from_action(head, eps, a_gpu, A):
    """log-prob with the tanh correction evaluated in fp32 from the kernel's own action: near |a| = 1 the term
    log(1 - a^2 + 1e-6) is ill-conditioned in fp32 (for the reference as well), so an fp64 tanh is no oracle there."""
    mu, raw = head[:, :A].double(), head[:, A:].double()
    ls = raw.clamp(-5, 2)
    e = eps.double() if eps is not None else torch.zeros_like(mu)
    lp = (-(e ** 2) / 2 - ls - 0.5 * math.log(2 * math.pi)).sum(-1)
    a32 = a_gpu.float().cpu()
    return lp - torch.log((1 - a32 * a32) + 1e-6).double().sum(-1)


def test_tanh_gauss_sample_and_bwd(rt):
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(2)
    A, O, B, rep = 6, 17, 32, 4
    M = B * rep
    head = torch.randn(2 * B, 2 * A, generator=gen) * 2.5      # exercises the clamp on both sides
    eps = torch.randn(M, A, generator=gen)
    obs = torch.randn(B, O, generator=gen)
    hd, ed, od = head.to(DEV), eps.to(DEV), obs.to(DEV)
    X = torch.zeros(M, O + A, device=DEV)
    lp = torch.zeros(M, device=DEV)
    L.call("orlk_tanh_gauss_sample", hd.data_ptr(), 2 * A, B, rep, ed.data_ptr(), M, A, X.data_ptr() + 4 * O, O + A,
           lp.data_ptr(), od.data_ptr(), O, O, X.data_ptr(), O + A, rt.cur)
    hrep = head[B:].repeat_interleave(rep, 0)
    a_ref, _ = _head_ref(hrep.double(), eps.double(), A)
    _close(X[:, O:], a_ref, rtol=1e-5, atol=1e-6, msg="sampled action")
    _close(lp, _lp_from_action(hrep, eps, X[:, O:], A), rtol=1e-5, atol=1e-4, msg="log-prob")
    assert torch.equal(X[:, :O].cpu(), obs.repeat_interleave(rep, 0))
    # mode (eps == NULL)
    L.call("orlk_tanh_gauss_sample", hd.data_ptr(), 2 * A, B, rep, None, M, A, X.data_ptr() + 4 * O, O + A,
           lp.data_ptr(), None, 0, 0, None, 0, rt.cur)
    a_ref, _ = _head_ref(hrep.double(), None, A)
    _close(X[:, O:], a_ref, msg="mode action")
    _close(lp, _lp_from_action(hrep, None, X[:, O:], A), rtol=1e-5, atol=1e-4, msg="mode log-prob")
    # backward against autograd (rep = 1); pre-tanh values kept moderate so that fp32 and fp64 agree,
    # raw log-sigma still crosses both clamp bounds (gradient gating)
    h0 = torch.randn(B, 2 * A, generator=gen)
    h0[:, :A] *= 0.5
    h0[:, A:] *= 2.5
    head1 = h0.double().requires_grad_(True)
    eps1 = torch.randn(B, A, generator=gen) * 0.5 * torch.exp(-h0[:, A:].clamp(-5, 2))
    a1, lp1 = _head_ref(head1, eps1.double(), A)
    dA0, dA1, glp = torch.randn(B, A, generator=gen), torch.randn(B, A, generator=gen), torch.randn(B, generator=gen)
    loss = (a1 * (dA0 + dA1).double()).sum() + (lp1[:, 0] * glp.double()).sum()
    loss.backward()
    h1d, e1d = head1.detach().float().to(DEV), eps1.to(DEV)
    act = torch.zeros(B, A, device=DEV)
    lpd = torch.zeros(B, device=DEV)
    L.call("orlk_tanh_gauss_sample", h1d.data_ptr(), 2 * A, 0, 1, e1d.data_ptr(), B, A, act.data_ptr(), A, lpd.data_ptr(),
           None, 0, 0, None, 0, rt.cur)
    _close(lpd, lp1[:, 0], rtol=1e-5, atol=1e-4, msg="log-prob (moderate)")
    dh = torch.zeros(B, 2 * A, device=DEV)
    d01, gl = torch.stack([dA0, dA1]).to(DEV), glp.to(DEV)
    L.call("orlk_tanh_gauss_bwd", h1d.data_ptr(), 2 * A, e1d.data_ptr(), act.data_ptr(), A, d01.data_ptr(), 2, B * A,
           A, gl.data_ptr(), B, A, dh.data_ptr(), 2 * A, rt.cur)
    _close(dh, head1.grad, rtol=2e-4, atol=2e-4, msg="head backward")


def test_philox_fill_statistics(rt):
    from offlinerlkit_b200 import _lib as L
    n_n, n_u = 400_003, 100_001
    out = torch.zeros(n_n + n_u, device=DEV)
    ctr = torch.zeros(1, dtype=torch.int64, device=DEV)
    en = torch.ones(1, dtype=torch.int32, device=DEV)
    L.call("orlk_philox_fill", out.data_ptr(), n_n, n_u, -1.0, 1.0, 1234, ctr.data_ptr(), en.data_ptr(), rt.cur)
    nrm, uni = out[:n_n].double().cpu(), out[n_n:].double().cpu()
    assert abs(nrm.mean()) < 0.01 and abs(nrm.std() - 1) < 0.01
    assert abs((nrm ** 4).mean() - 3) < 0.1                      # kurtosis of a Gaussian
    assert uni.min() >= -1 and uni.max() < 1 and abs(uni.mean()) < 0.01 and abs(uni.var() - 1 / 3) < 0.01
    first = out.clone()
    L.call("orlk_philox_fill", out.data_ptr(), n_n, n_u, -1.0, 1.0, 1234, ctr.data_ptr(), en.data_ptr(), rt.cur)
    assert torch.equal(first, out)                               # same counter -> same stream (reproducible)
    ctr += 1
    L.call("orlk_philox_fill", out.data_ptr(), n_n, n_u, -1.0, 1.0, 1234, ctr.data_ptr(), en.data_ptr(), rt.cur)
    assert not torch.equal(first, out)
    en.zero_()
    out.fill_(7.0)
    L.call("orlk_philox_fill", out.data_ptr(), n_n, n_u, -1.0, 1.0, 1234, ctr.data_ptr(), en.data_ptr(), rt.cur)
    assert (out == 7.0).all()                                    # disabled: injected noise is left untouched


# ------------------------------------------------------------------------------------------------ optimiser
def test_adam_polyak_matches_torch(rt):
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import AdamT
    gen = torch.Generator().manual_seed(3)
    n, splits = 5000, 3
    p0 = torch.randn(n, generator=gen)
    tgt0 = torch.randn(n, generator=gen)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=3e-4)
    pd, md, vd, td = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), tgt0.to(DEV)
    groups = (L.AdamGroup * 2)()
    groups[1].lr, groups[1].beta1, groups[1].beta2, groups[1].eps, groups[1].tau, groups[1].step = 3e-4, 0.9, 0.999, 1e-8, 0.005, 0
    groups[1].refresh()
    gd = torch.frombuffer(bytearray(bytes(groups)), dtype=torch.uint8).to(DEV)
    gbuf = torch.zeros(splits, n, device=DEV)
    tref = tgt0.clone()
    op = rt.adam([AdamT(p=pd.data_ptr(), n=n, group=1, m=md.data_ptr(), v=vd.data_ptr(), tgt=td.data_ptr(),
                        grad=gbuf.data_ptr(), g_splits=splits, g_split_stride=n, flags=L.OPT_ADAM | L.OPT_POLYAK)],
                 gd.data_ptr())
    for it in range(5):
        parts = torch.randn(splits, n, generator=gen) * (10.0 ** (it - 3))
        gbuf.copy_(parts)
        pt.grad = parts.sum(0)
        opt.step()
        tref = tref * (1 - 0.005) + pt.detach() * 0.005
        op()
        L.call("orlk_step_end", gd.data_ptr(), 0b10, None, rt.cur)
        torch.cuda.synchronize()
        _close(pd, pt.detach(), rtol=1e-6, atol=2e-7, msg=f"adam param it{it}")
        _close(td, tref, rtol=1e-6, atol=2e-7, msg=f"polyak it{it}")


# ------------------------------------------------------------------------------------------------ losses
def test_sac_actor_loss_vs_autograd(rt):
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(4)
    for (E, B, clamp) in [(2, 256, 0), (2, 16, 1), (10, 256, 1)]:
        q = torch.randn(E, B, generator=gen)
        q[0, :3] = q[1, :3]                                     # ties
        logp = torch.randn(B, generator=gen)
        la0, H = 0.3, -6.0
        qd = q.clone().double().requires_grad_(True)
        lpd = logp.clone().double().requires_grad_(True)
        alpha = math.exp(la0)
        mn = torch.min(qd[0], qd[1]) if E == 2 else torch.min(qd, 0)[0]
        loss = (alpha * lpd - mn).mean()
        loss.backward()
        la = torch.tensor([la0], requires_grad=True)
        aopt = torch.optim.Adam([la], lr=1e-2)
        aloss = -(la * (logp + H)).mean()
        aloss.backward()
        aopt.step()
        new_alpha = la.detach().exp().clamp(0, 1) if clamp else la.detach().exp()
        sc = torch.zeros(8, device=DEV)
        sc[0], sc[1] = la0, alpha
        groups = (L.AdamGroup * 1)()
        groups[0].lr, groups[0].beta1, groups[0].beta2, groups[0].eps = 1e-2, 0.9, 0.999, 1e-8
        groups[0].refresh()
        gd = torch.frombuffer(bytearray(bytes(groups)), dtype=torch.uint8).to(DEV)
        mv = torch.zeros(2, device=DEV)
        qg, lg = q.to(DEV), logp.to(DEV)
        dq, glp, out = torch.zeros(E, B, device=DEV), torch.zeros(B, device=DEV), torch.zeros(4, device=DEV)
        L.call("orlk_sac_actor_loss", qg.data_ptr(), B, E, lg.data_ptr(), B, sc.data_ptr(), 1, clamp, H, gd.data_ptr(), 0,
               mv.data_ptr(), dq.data_ptr(), B, glp.data_ptr(), out.data_ptr(), rt.cur)
        torch.cuda.synchronize()
        _close(out[0], loss, rtol=1e-5, atol=1e-5, msg="actor loss")
        _close(out[1], aloss, rtol=1e-5, atol=1e-5, msg="alpha loss")
        _close(out[2], new_alpha[0], rtol=1e-5, msg="alpha")
        _close(sc[1], new_alpha[0], rtol=1e-5, msg="alpha scalar")
        _close(dq, qd.grad, rtol=1e-6, atol=1e-9, msg="dq")
        _close(glp, lpd.grad, rtol=1e-6, atol=1e-9, msg="dlogp")


def test_cql_critic_loss_vs_autograd(rt):
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(6)
    # (B, rows feeding the conservative term, rows in the `- w mean Q` term): CQL has all three equal; COMBO's
    # rho_s="mix" has nq = real rows, rho_s="model" additionally draws the conservative rows from the fake half
    # the last field: max_q_backup (cql.py:109-120) -- N target values per row, max over them, no entropy term
    for (B, N, A, det, lag, Bc, nq_rows, mqb) in [(16, 4, 3, 1, 0, 16, 16, 0), (256, 10, 6, 0, 1, 256, 256, 0),
                                                  (256, 10, 6, 1, 1, 256, 256, 0), (256, 10, 6, 1, 1, 256, 128, 0),
                                                  (256, 10, 6, 0, 0, 128, 128, 0), (24, 4, 3, 1, 1, 15, 9, 0),
                                                  (16, 4, 3, 0, 0, 16, 16, 1), (256, 10, 6, 1, 1, 256, 128, 1)]:
        R = Bc * N
        Mc = B + 3 * R
        rep = N if mqb else 1
        q = torch.randn(2, Mc, generator=gen) * 3
        tq = torch.randn(2, B * rep, generator=gen)
        lpn, lpp, lpq = torch.randn(B, generator=gen), torch.randn(R, generator=gen), torch.randn(R, generator=gen)
        rew, term = torch.randn(B, generator=gen), (torch.rand(B, generator=gen) < 0.1).float()
        gamma, w, T, thr, alpha, cla0 = 0.99, 5.0, 1.3, 10.0, 0.7, 0.2
        qd = q.clone().double().requires_grad_(True)
        if mqb:
            nq = torch.min(tq[0].view(B, N).max(1)[0], tq[1].view(B, N).max(1)[0]).double()
        else:
            nq = torch.min(tq[0], tq[1]).double()
            if not det:
                nq = nq - alpha * lpn.double()
        y = rew.double() + gamma * (1 - term.double()) * nq
        cla = torch.tensor([cla0], dtype=torch.double, requires_grad=True)
        losses = []
        for c in range(2):
            td = ((qd[c, :B] - y) ** 2).mean()
            cat = torch.stack([qd[c, B:B + R] - lpp.double(), qd[c, B + R:B + 2 * R] - lpq.double(),
                               qd[c, B + 2 * R:] - math.log(0.5 ** A)], 1)
            cons = torch.logsumexp(cat / T, dim=1).mean() * w * T - qd[c, :nq_rows].mean() * w
            if lag:
                cons = torch.clamp(cla.exp(), 0, 1e6) * (cons - thr)
            losses.append((td, cons))
        cql_alpha_loss = -(losses[0][1] + losses[1][1]) * 0.5 if lag else None
        if lag:
            g_cla, = torch.autograd.grad(cql_alpha_loss, cla, retain_graph=True)
        total = [td + cons for td, cons in losses]
        (total[0] + total[1]).backward()
        sc = torch.zeros(8, device=DEV)
        sc[1], sc[2] = alpha, cla0
        groups = (L.AdamGroup * 1)()
        groups[0].lr, groups[0].beta1, groups[0].beta2, groups[0].eps = 3e-4, 0.9, 0.999, 1e-8
        groups[0].refresh()
        gd = torch.frombuffer(bytearray(bytes(groups)), dtype=torch.uint8).to(DEV)
        mv = torch.zeros(2, device=DEV)
        dev = lambda t: t.to(DEV)
        qg, tqg, a1, a2, a3, rg, tg = map(dev, (q, tq, lpn, lpp, lpq, rew, term))
        dq, out = torch.zeros(2, Mc, device=DEV), torch.zeros(4, device=DEV)
        scratch = torch.zeros(L.load().orlk_cql_critic_loss_scratch_floats(B, R), device=DEV)
        L.call("orlk_cql_critic_loss", qg.data_ptr(), Mc, tqg.data_ptr(), B * rep, a1.data_ptr(), a2.data_ptr(), a3.data_ptr(),
               rg.data_ptr(), tg.data_ptr(), B, nq_rows, rep, R, A, gamma, w, T, det, lag, thr, sc.data_ptr(), gd.data_ptr(), 0,
               mv.data_ptr(), dq.data_ptr(), Mc, out.data_ptr(), scratch.data_ptr(), rt.cur)
        torch.cuda.synchronize()
        assert scratch.view(torch.int32)[0].item() == 0, "the block counter must be left at zero for the next launch"
        _close(out[0], total[0], rtol=2e-5, atol=1e-4, msg="critic1 loss")
        _close(out[1], total[1], rtol=2e-5, atol=1e-4, msg="critic2 loss")
        _close(dq, qd.grad, rtol=2e-5, atol=1e-8, msg="dq")
        if lag:
            _close(out[2], cql_alpha_loss, rtol=2e-5, atol=1e-4, msg="cql alpha loss")
            _close(out[3], cla.detach().exp()[0], rtol=1e-5, msg="cql alpha")
            # first Adam step moves log-alpha by -lr*sign(g)
            _close(sc[2], torch.tensor(cla0 - 3e-4 * math.copysign(1.0, g_cla.item())), rtol=1e-4, msg="cql log alpha")


# ------------------------------------------------------------------------------------------------ replay
def test_replay_gather_bit_exact_and_ring(rt):
    from offlinerlkit_b200.buffer import ReplayBuffer
    from offlinerlkit_b200.synthetic import make_dataset
    for (O, A, n) in [(11, 3, 5000), (17, 6, 3001)]:
        d = make_dataset(n, O, A, seed=1)
        buf = ReplayBuffer(n, (O,), np.float32, A, np.float32, device=DEV)
        buf.load_dataset(d)
        for B in (1, 256, 1000):
            np.random.seed(B)
            batch = buf.sample(B)
            np.random.seed(B)
            idx = np.random.randint(0, n, size=B)
            torch.cuda.synchronize()
            assert np.array_equal(batch.indices.cpu().numpy(), idx)
            assert np.array_equal(batch["observations"].cpu().numpy(), d["observations"][idx])
            assert np.array_equal(batch["next_observations"].cpu().numpy(), d["next_observations"][idx])
            assert np.array_equal(batch["actions"].cpu().numpy(), d["actions"][idx])
            assert np.array_equal(batch["rewards"].cpu().numpy(), d["rewards"][idx].reshape(-1, 1))
            assert np.array_equal(batch["terminals"].cpu().numpy(), d["terminals"][idx].reshape(-1, 1))
    # ring writes (add_batch wrap-around) are mirrored before the next sample
    buf = ReplayBuffer(100, (4,), np.float32, 2, np.float32, device=DEV)
    rng = np.random.default_rng(0)
    for k in range(5):
        m = 37
        buf.add_batch(rng.standard_normal((m, 4), dtype=np.float32), rng.standard_normal((m, 4), dtype=np.float32),
                      rng.standard_normal((m, 2), dtype=np.float32), rng.standard_normal((m, 1), dtype=np.float32),
                      (rng.random((m, 1)) < 0.5).astype(np.float32))
        idx = np.arange(buf._size)
        b = buf.gather(idx)
        torch.cuda.synchronize()
        assert np.array_equal(b["observations"].cpu().numpy(), buf.observations[idx])
        assert np.array_equal(b["terminals"].cpu().numpy(), buf.terminals[idx])
    mean, std = buf.normalize_obs()
    b = buf.gather(np.arange(buf._size))
    torch.cuda.synchronize()
    assert np.array_equal(b["next_observations"].cpu().numpy(), buf.next_observations[:buf._size])


def test_replay_add_batch_from_device_tensors(rt):
    """Rollout hand-off (mb_policy_trainer.py:71-73) with CUDA tensors: the host arrays get the reference's contents
    (same ring arithmetic, incl. wrap-around and uint8 terminals) and the device table is written directly - a sample
    taken afterwards returns exactly the host rows, without a re-upload."""
    from offlinerlkit_b200.buffer import ReplayBuffer
    ref = ReplayBuffer(100, (4,), np.float32, 2, np.float32, device=DEV)       # fed with NumPy arrays
    buf = ReplayBuffer(100, (4,), np.float32, 2, np.float32, device=DEV)       # fed with CUDA tensors
    rng = np.random.default_rng(5)
    lazy = ReplayBuffer(100, (4,), np.float32, 2, np.float32, device=DEV)      # CUDA tensors, host arrays read only at the end
    lazy.add_batch(np.zeros((1, 4), np.float32), np.zeros((1, 4), np.float32), np.zeros((1, 2), np.float32),
                   np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32))
    lazy.gather(np.arange(1))                                                   # creates the device table
    ref2 = ReplayBuffer(100, (4,), np.float32, 2, np.float32, device=DEV)
    ref2.add_batch(np.zeros((1, 4), np.float32), np.zeros((1, 4), np.float32), np.zeros((1, 2), np.float32),
                   np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32))
    for k, m in enumerate((37, 37, 37, 12, 99, 100, 130)):
        parts = (rng.standard_normal((m, 4), dtype=np.float32), rng.standard_normal((m, 4), dtype=np.float32),
                 rng.standard_normal((m, 2), dtype=np.float32), rng.standard_normal((m, 1), dtype=np.float32),
                 (rng.random((m, 1)) < 0.5))
        ref.add_batch(*[p.astype(np.float32) for p in parts])
        dev = [torch.from_numpy(p).to(DEV) for p in parts[:4]] + [torch.from_numpy(parts[4].astype(np.uint8)).to(DEV)]
        lazy.add_batch(*dev)
        ref2.add_batch(*[p.astype(np.float32) for p in parts])
        assert lazy._host_pending_rows < 2 * 100      # bounded backlog: covered segments are dropped
        buf.add_batch(*dev)
        if k >= 1 and m < 100:
            assert not buf._dirty, "device batches must not schedule a host -> device re-upload"
        assert buf._ptr == ref._ptr and buf._size == ref._size
        for name in ("observations", "next_observations", "actions", "rewards", "terminals"):
            assert np.array_equal(getattr(buf, name), getattr(ref, name)), (k, name)
        idx = np.arange(buf._size)
        b = buf.gather(idx)
        torch.cuda.synchronize()
        for name in ("observations", "next_observations", "actions", "rewards", "terminals"):
            assert np.array_equal(b[name].cpu().numpy(), getattr(ref, name)[idx]), (k, name)
    # the lazily mirrored buffer: several ring laps without a single host read, then everything must be there
    for name in ("observations", "next_observations", "actions", "rewards", "terminals"):
        assert np.array_equal(getattr(lazy, name), getattr(ref2, name)), name
    assert not lazy._host_pending


# ------------------------------------------------------------------------------------------------ tcgen05 GEMM
def _tc_case(rt, G, M, N, K, passes, epi, splits, with_bias, want_ct, want_rowsum, seed, n_tile=0, a_mn=False, b_mn=False,
             gen_mode=False):
    """gen_mode: the kernel multiplies A' = row[m] * col[k] * (A > 0) instead of A (the fused scalar-head backward)."""
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import Mat
    gen = torch.Generator().manual_seed(seed)
    A = torch.randn(G, M, K, generator=gen)
    B = torch.randn(G, N, K, generator=gen) / math.sqrt(K)
    bias = torch.randn(G, N, generator=gen)
    aux = torch.randn(G, M, N, generator=gen)
    grow, gcol = torch.randn(G, M, generator=gen), torch.randn(G, K, generator=gen)
    Ad, Bd, bd, auxd = A.to(DEV), B.to(DEV), bias.to(DEV), aux.to(DEV)
    growd, gcold = grow.to(DEV), gcol.to(DEV)
    if a_mn:
        Ad = Ad.transpose(1, 2).contiguous()        # stored [G][K][M]
    if b_mn:
        Bd = Bd.transpose(1, 2).contiguous()        # stored [G][K][N]
    Amat = Mat(Ad.data_ptr(), K, M, M) if a_mn else Mat(Ad.data_ptr(), M, K, K)
    Bmat = Mat(Bd.data_ptr(), K, N, N) if b_mn else Mat(Bd.data_ptr(), N, K, K)
    gkw = dict(gen_row=growd.data_ptr(), gen_row_gs=M, gen_col=gcold.data_ptr(), gen_col_gs=K) if gen_mode else {}
    if gen_mode:
        A = grow[:, :, None] * gcol[:, None, :] * (A > 0)
    s_eff = rt.lib.orlk_tc_effective_splits(K, splits)
    Cd = torch.full((s_eff, G, M, N), float("nan"), device=DEV)
    CTd = torch.full((G, N, M), float("nan"), device=DEV)
    rs = torch.full((s_eff, G, M), float("nan"), device=DEV)
    op = rt.tc_gemm(A=Amat, a_gs=M * K, B=Bmat, b_gs=N * K, G=G, a_mn=a_mn, b_mn=b_mn, **gkw,
                    passes=passes, epi=epi, C=Mat(Cd.data_ptr(), M, N, N), c_gs=M * N, c_split_stride=G * M * N,
                    CT=Mat(CTd.data_ptr(), N, M, M) if want_ct else None, ct_gs=N * M,
                    bias=bd.data_ptr() if with_bias else 0, bias_gs=N,
                    aux=Mat(auxd.data_ptr(), M, N, N) if epi == L.EPI_RELU_MASK else None, aux_gs=M * N,
                    rowsum=rs.data_ptr() if want_rowsum else 0, rowsum_gs=M, rowsum_split_stride=G * M, k_splits=splits,
                    n_tile=n_tile)
    op()
    torch.cuda.synchronize()
    ref = torch.einsum("gmk,gnk->gmn", A.double(), B.double())
    if with_bias:
        ref = ref + bias.double()[:, None, :]
    if epi == L.EPI_RELU:
        ref = ref.clamp(min=0)
    elif epi == L.EPI_RELU_MASK:
        ref = ref * (aux > 0)
    got = Cd.sum(0) if s_eff > 1 else Cd[0]
    tag = f"tc G{G} {M}x{N}x{K} passes{passes} epi{epi} splits{s_eff}"
    assert not torch.isnan(got).any(), tag + ": unwritten output"
    # 3 passes: fp32-grade.  1 pass: TF32 operands (10-bit mantissa, truncated) -> ~1e-3 of the row scale
    tol = 3e-6 if passes == 3 else 3e-3
    scale = (A.double().abs() @ B.double().abs().transpose(1, 2)).max().item()
    err = (got.double().cpu() - ref).abs().max().item()
    assert err <= tol * scale, f"{tag}: max err {err:.3e} vs {tol * scale:.3e}"
    if want_ct:
        assert torch.equal(CTd, got.transpose(1, 2).contiguous()) or s_eff > 1, tag + " CT"
    if want_rowsum:
        rref = A.double().sum(2)
        rerr = (rs.sum(0).double().cpu() - rref).abs().max().item()
        assert rerr <= (3e-6 if passes == 3 else 3e-3) * A.abs().sum(2).max().item(), f"{tag} rowsum err {rerr:.3e}"
    return err / scale


@pytest.mark.parametrize("passes", [3, 1])
def test_tc_gemm_forward_dgrad_wgrad_shapes(rt, passes):
    from offlinerlkit_b200 import _lib as L
    errs = []
    errs.append(_tc_case(rt, 1, 128, 256, 32, passes, L.EPI_NONE, 1, False, False, False, 1))     # one slab, one tile
    errs.append(_tc_case(rt, 1, 128, 256, 256, passes, L.EPI_NONE, 1, False, True, False, 2))
    errs.append(_tc_case(rt, 2, 300, 256, 256, passes, L.EPI_RELU, 1, True, True, False, 3))       # ragged M, bias+ReLU
    errs.append(_tc_case(rt, 2, 7936, 256, 256, passes, L.EPI_RELU_MASK, 1, False, True, False, 4))  # critic dgrad
    errs.append(_tc_case(rt, 2, 256, 256, 7936, passes, L.EPI_NONE, 16, False, False, True, 5))    # critic wgrad + bias grads
    errs.append(_tc_case(rt, 3, 200, 64, 96, passes, L.EPI_NONE, 1, True, True, True, 6))          # narrow N, K not /128
    errs.append(_tc_case(rt, 1, 1000, 208, 224, passes, L.EPI_RELU, 1, True, False, False, 7))     # N = 208 (13 x 16)
    # small-M layers: the output columns are tiled 32 per CTA
    errs.append(_tc_case(rt, 2, 256, 256, 256, passes, L.EPI_RELU, 1, True, True, False, 8, n_tile=32))
    errs.append(_tc_case(rt, 2, 512, 256, 256, passes, L.EPI_RELU_MASK, 1, False, True, False, 9, n_tile=32))
    errs.append(_tc_case(rt, 1, 200, 192, 64, passes, L.EPI_NONE, 1, True, False, True, 10, n_tile=64))
    # MN-major operands (the weight gradient reads row-major activations / gradients as they are)
    errs.append(_tc_case(rt, 2, 256, 256, 7936, passes, L.EPI_NONE, 31, False, False, True, 11, a_mn=True, b_mn=True))
    errs.append(_tc_case(rt, 1, 200, 96, 300, passes, L.EPI_NONE, 1, False, False, True, 12, a_mn=True, b_mn=True, n_tile=32))
    errs.append(_tc_case(rt, 2, 300, 256, 256, passes, L.EPI_RELU, 1, True, False, False, 13, a_mn=True))
    errs.append(_tc_case(rt, 2, 300, 256, 256, passes, L.EPI_NONE, 1, False, False, False, 14, b_mn=True))
    # ensembles of short members (EDAC critics, 'io' weights): forward with an MN-major B in 64-column tiles, 50 members
    # in full-width tiles (several waves), the input-gradient shape (K-major B, ReLU mask) and the weight gradient dW[i][o]
    errs.append(_tc_case(rt, 10, 256, 256, 256, passes, L.EPI_RELU, 1, True, False, False, 42, n_tile=64, b_mn=True))
    errs.append(_tc_case(rt, 50, 256, 256, 256, passes, L.EPI_RELU, 1, True, False, False, 43, n_tile=256, b_mn=True))
    errs.append(_tc_case(rt, 10, 256, 256, 256, passes, L.EPI_RELU_MASK, 1, False, False, False, 44, n_tile=64))
    errs.append(_tc_case(rt, 10, 256, 256, 256, passes, L.EPI_NONE, 1, False, False, False, 45, n_tile=64, a_mn=True, b_mn=True))
    # rank-1 operand generator (scalar-head backward folded into dgrad / wgrad), both operand orders
    errs.append(_tc_case(rt, 2, 7936, 256, 256, passes, L.EPI_RELU_MASK, 1, False, True, False, 15, gen_mode=True))
    errs.append(_tc_case(rt, 2, 256, 256, 7936, passes, L.EPI_NONE, 31, False, False, True, 16, gen_mode=True))
    errs.append(_tc_case(rt, 2, 256, 256, 7936, passes, L.EPI_NONE, 31, False, False, True, 17, gen_mode=True, a_mn=True,
                         b_mn=True))
    # first-layer weight gradient: N = 23 columns inside a 32-column tile, operands MN-major, B shared by the groups
    errs.append(_tc_wgrad0_case(rt, passes))
    # first-layer shape: K = 23 with unaligned weight rows (staged by the kernel's own warps), A shared by the groups
    errs.append(_tc_first_layer_case(rt, passes))
    print(f"passes={passes}: relative errors {['%.2e' % e for e in errs]}")


def _tc_wgrad0_case(rt, passes):
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import Mat
    gen = torch.Generator().manual_seed(78)
    G, Mb, O, I = 2, 7936, 256, 23
    dZ = torch.randn(G, Mb, O, generator=gen) / math.sqrt(Mb)
    X = torch.randn(Mb, I, generator=gen)
    Xd = torch.zeros(Mb, 24, device=DEV)
    Xd[:, :I] = X.to(DEV)
    dZd = dZ.to(DEV)
    s_eff = rt.lib.orlk_tc_effective_splits(Mb, 31)
    Cd = torch.full((s_eff, G, O, I), float("nan"), device=DEV)
    rs = torch.full((s_eff, G, O), float("nan"), device=DEV)
    rt.tc_gemm(A=Mat(dZd.data_ptr(), Mb, O, O), a_gs=Mb * O, a_mn=True, B=Mat(Xd.data_ptr(), Mb, I, 24), b_gs=0, b_mn=True,
               n_tile=32, G=G, passes=passes, C=Mat(Cd.data_ptr(), O, I, I), c_gs=O * I, c_split_stride=G * O * I,
               rowsum=rs.data_ptr(), rowsum_gs=O, rowsum_split_stride=G * O, k_splits=31)()
    torch.cuda.synchronize()
    ref = torch.einsum("gmo,mi->goi", dZ.double(), X.double())
    got = Cd.sum(0).double().cpu()
    assert not torch.isnan(got).any(), "first-layer wgrad: unwritten output"
    scale = torch.einsum("gmo,mi->goi", dZ.double().abs(), X.double().abs()).max().item()
    err = (got - ref).abs().max().item()
    assert err <= (3e-6 if passes == 3 else 3e-3) * scale, f"first-layer wgrad on tensor cores: err {err:.3e} scale {scale:.3e}"
    rerr = (rs.sum(0).double().cpu() - dZ.double().sum(1)).abs().max().item()
    assert rerr <= (3e-6 if passes == 3 else 3e-3) * dZ.abs().sum(1).max().item(), f"first-layer wgrad bias sums: {rerr:.3e}"
    return err / scale


def _tc_first_layer_case(rt, passes):
    from offlinerlkit_b200 import _lib as L
    from offlinerlkit_b200.engine.core import Mat
    gen = torch.Generator().manual_seed(77)
    G, M, N, K = 2, 1000, 256, 23
    X = torch.randn(M, K, generator=gen)
    W = torch.randn(G, N, K, generator=gen) / math.sqrt(K)
    b = torch.randn(G, N, generator=gen)
    Xd = torch.zeros(M, 24, device=DEV)
    Xd[:, :K] = X.to(DEV)
    Wd, bd = W.to(DEV), b.to(DEV)
    Cd = torch.full((G, M, N), float("nan"), device=DEV)
    rt.tc_gemm(A=Mat(Xd.data_ptr(), M, K, 24), a_gs=0, B=Mat(Wd.data_ptr(), N, K, K), b_gs=N * K, G=G, passes=passes,
               epi=L.EPI_RELU, C=Mat(Cd.data_ptr(), M, N, N), c_gs=M * N, bias=bd.data_ptr(), bias_gs=N)()
    torch.cuda.synchronize()
    ref = (torch.einsum("mk,gnk->gmn", X.double(), W.double()) + b.double()[:, None, :]).clamp(min=0)
    scale = (X.double().abs() @ W.double().abs().transpose(1, 2)).max().item()
    err = (Cd.double().cpu() - ref).abs().max().item()
    assert err <= (3e-6 if passes == 3 else 3e-3) * scale, f"first layer on tensor cores: err {err:.3e}"
    return err / scale


# ------------------------------------------------------------------------------------------------ narrow layers (large M)
def test_narrow_fwd_and_wgrad(rt):
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(9)
    for (G, M, N, K) in [(2, 7936, 256, 23), (1, 1000, 200, 14), (3, 333, 40, 32)]:
        X = torch.randn(M, K, generator=gen)
        W = torch.randn(G, N, K, generator=gen)
        b = torch.randn(G, N, generator=gen)
        Xd, Wd, bd = X.to(DEV), W.to(DEV), b.to(DEV)
        Mt = (M + 3) // 4 * 4
        Y, YT = torch.zeros(G, M, N, device=DEV), torch.zeros(G, N, Mt, device=DEV)
        L.call("orlk_narrow_fwd", Xd.data_ptr(), K, 0, Wd.data_ptr(), K, N * K, bd.data_ptr(), N, Y.data_ptr(), N, M * N,
               YT.data_ptr(), Mt, N * Mt, M, N, K, G, 1, rt.cur)
        ref = (torch.einsum("mk,gnk->gmn", X.double(), W.double()) + b.double()[:, None, :]).clamp(min=0)
        _close(Y, ref, msg=f"narrow fwd {G,M,N,K}")
        assert torch.equal(YT[:, :, :M], Y.transpose(1, 2).contiguous()), "narrow fwd transposed copy"
        # first-layer weight gradient: dW[g][n][k] = sum_m dZ[g][m][n] X[m][k], db[g][n] = sum_m dZ[g][m][n]
        dZ = torch.randn(G, M, N, generator=gen)
        dZd = dZ.to(DEV)
        ch = rt.lib.orlk_narrow_wgrad_chunks(M)
        wp, bp = torch.zeros(ch, G, N, K, device=DEV), torch.zeros(ch, G, N, device=DEV)
        L.call("orlk_narrow_wgrad", dZd.data_ptr(), N, M * N, Xd.data_ptr(), K, 0, wp.data_ptr(), 1, K, N * K, G * N * K,
               bp.data_ptr(), N, G * N, None, 0, 0, M, N, K, G, rt.cur)
        _close(wp.sum(0), torch.einsum("gmn,mk->gnk", dZ.double(), X.double()), rtol=1e-5, atol=1e-4, msg="narrow wgrad W0")
        _close(bp.sum(0), dZ.double().sum(1), rtol=1e-5, atol=1e-4, msg="narrow wgrad b0")
    # head weight gradient: dW[g][ns][k] = sum_m dOut[g][m][ns] H[g][m][k], db[g][ns] = sum_m dOut[g][m][ns]
    for (G, M, K, NS) in [(2, 7936, 256, 1), (1, 2000, 256, 12)]:
        H, dO = torch.randn(G, M, K, generator=gen), torch.randn(G, M, NS, generator=gen)
        Hd, dOd = H.to(DEV), dO.to(DEV)
        ch = rt.lib.orlk_narrow_wgrad_chunks(M)
        wp, bp = torch.zeros(ch, G, NS, K, device=DEV), torch.zeros(ch, G, NS, device=DEV)
        L.call("orlk_narrow_wgrad", Hd.data_ptr(), K, M * K, dOd.data_ptr(), NS, M * NS, wp.data_ptr(), K, 1, NS * K,
               G * NS * K, None, 0, 0, bp.data_ptr(), NS, G * NS, M, K, NS, G, rt.cur)
        _close(wp.sum(0), torch.einsum("gmn,gmk->gnk", dO.double(), H.double()), rtol=1e-5, atol=1e-4, msg="head wgrad W")
        _close(bp.sum(0), dO.double().sum(1), rtol=1e-5, atol=1e-4, msg="head wgrad b")


# ------------------------------------------------------------------------------------------------ rollout compaction
@pytest.mark.parametrize("S", [1, 100, 4096, 5000, 50_000, 70_001])
@pytest.mark.parametrize("p_drop", [0.0, 0.3, 1.0])
def test_compact_rows_is_stable_and_exact(rt, S, p_drop):
    """mopo.py:69-73 (`observations = next_observations[nonterm_mask]`): survivors keep their order, rows are copied bit
    for bit, the count is the number of survivors.  Small inputs take the single-CTA kernel, large ones the two-launch
    block-scan form."""
    from offlinerlkit_b200 import _lib as L
    gen = torch.Generator().manual_seed(S)
    w, ld = 17, 20
    src = torch.randn(S, ld, generator=gen).to(DEV)
    drop = (torch.rand(S, generator=gen) < p_drop).to(torch.uint8).to(DEV)
    dst = torch.full((S, w), float("nan"), device=DEV)
    count = torch.full((1,), -1, dtype=torch.int32, device=DEV)
    L.call("orlk_compact_rows", drop.data_ptr(), S, src.data_ptr(), ld, w, dst.data_ptr(), w, count.data_ptr(), rt.cur)
    torch.cuda.synchronize()
    want = src[:, :w][drop == 0]
    assert int(count.item()) == want.shape[0]
    assert torch.equal(dst[:want.shape[0]], want)


# ------------------------------------------------------------------------------------------------ fused critic passes
def _fused_case(rt, M, N, K0, nh, G, seed, store_h=True):
    """A random Linear+ReLU stack laid out like a ParamSet (one 4-float aligned block per member) + the fused-forward job."""
    from offlinerlkit_b200.engine.core import Mat
    gen = torch.Generator().manual_seed(seed)
    ldx = (K0 + 3) // 4 * 4
    X = torch.zeros(M, ldx)
    X[:, :K0] = torch.randn(M, K0, generator=gen)

    def al(n):
        return (n + 3) // 4 * 4
    offs, off = [], 0
    dims = [(N, K0)] + [(N, N)] * (nh - 1) + [(1, N)]
    for (o, i) in dims:
        w_off = off; off = al(off + o * i)
        b_off = off; off = al(off + o)
        offs.append((w_off, b_off))
    block = al(off)
    P = torch.zeros(G * block)
    Ws, bs = [], []
    for g in range(G):
        Wg, bg = [], []
        for (o, i), (wo, bo) in zip(dims, offs):
            W = torch.randn(o, i, generator=gen) / math.sqrt(i)
            b = torch.randn(o, generator=gen) * 0.3
            P[g * block + wo:g * block + wo + o * i] = W.reshape(-1)
            P[g * block + bo:g * block + bo + o] = b
            Wg.append(W); bg.append(b)
        Ws.append(Wg); bs.append(bg)
    Xd, Pd = X.to(DEV), P.to(DEV)
    Plo = torch.full_like(Pd, float("nan"))
    pad = torch.full((2, G, N, 32), float("nan"), device=DEV)
    base, lo = Pd.data_ptr(), Plo.data_ptr()
    rt.fused_prep(Pd, Plo, W0=base + 4 * offs[0][0], gs=block, N=N, K0=K0, G=G, w0pad=pad)()
    H = [torch.full((G, M, N), float("nan"), device=DEV) for _ in range(nh)] if store_h else None
    out = torch.full((G, M), float("nan"), device=DEV)
    job = rt.fused_fwd_job(X=Mat(Xd.data_ptr(), M, K0, ldx), W0pad=pad[0].data_ptr(), W0pad_lo=pad[1].data_ptr(),
                           W=[0] + [base + 4 * offs[l][0] for l in range(1, nh)],
                           Wlo=[0] + [lo + 4 * offs[l][0] for l in range(1, nh)], bias=[base + 4 * offs[l][1] for l in range(nh)],
                           H=[h.data_ptr() for h in H] if store_h else None, gs=block, h_gs=M * N,
                           head_w=base + 4 * offs[nh][0], head_b=base + 4 * offs[nh][1], out=out.data_ptr(), out_gs=M,
                           M=M, N=N, K0=K0, G=G)
    return dict(job=job, X=X, P=P, Ws=Ws, bs=bs, H=H, out=out, Plo=Plo, pad=pad, offs=offs, block=block, keep=(Xd, Pd),
                dims=(M, N, K0, nh, G))


def _check_fused(case):
    M, N, K0, nh, G = case["dims"]
    X, Ws, bs, H, out = case["X"], case["Ws"], case["bs"], case["H"], case["out"]
    for g in range(G):
        # 3xTF32: ~3e-6 of sum_k |a_k b_k| per layer (the tolerance of the per-layer tensor-core test); errors add up
        h = X[:, :K0].double()
        tol = 0.0
        for l in range(nh + 1):
            W, b = Ws[g][l].double(), bs[g][l].double()
            tol += 3e-6 * (h.abs() @ W.abs().t()).max().item()
            h = h @ W.t() + b
            if l < nh:
                h = torch.relu(h)
                if H is None:
                    continue
                got = H[l][g]
            else:
                got = out[g][:, None]
            assert not torch.isnan(got).any(), f"layer {l} member {g}: unwritten output"
            err = (got.double().cpu() - h).abs().max().item()
            assert err <= tol, f"layer {l} member {g}: max err {err:.3e} vs {tol:.3e}"


@pytest.mark.parametrize("pairs", [False, True])
@pytest.mark.parametrize("M,N,K0,nh,G", [(7936, 256, 23, 3, 2), (2560, 256, 14, 2, 2), (300, 128, 32, 4, 1),
                                         (2048 + 77, 64, 5, 3, 3)])
def test_critic_forward_fused_matches_fp64(rt, M, N, K0, nh, G, pairs):
    """orlk_critic_fwd_fused: every hidden activation and the scalar head of all members against an fp64 evaluation of
    the same Linear+ReLU stack (nets/mlp.py:22-28, modules/critic_module.py:25-33); 3xTF32 = fp32-grade.  Single CTAs and
    CTA pairs (cta_group::2; odd strip counts get a padding partner)."""
    case = _fused_case(rt, M, N, K0, nh, G, seed=M + N + K0)
    op = rt.critic_fwd_fused([case["job"]], pairs=pairs)
    for _ in range(2):          # a second launch over the same buffers: barriers / ring state start clean every time
        op()
    torch.cuda.synchronize()
    P = case["P"]
    lo_ref = P - (P.view(torch.int32) & ~0x1FFF).view(torch.float32)
    assert torch.equal(case["Plo"].cpu(), lo_ref), "lo words"
    w0, _ = case["offs"][0]
    for g in range(G):
        W0 = P[g * case["block"] + w0:g * case["block"] + w0 + N * K0].view(N, K0)
        padded = torch.zeros(N, 32)
        padded[:, :K0] = W0
        assert torch.equal(case["pad"][0, g].cpu(), padded), "padded first layer"
        assert torch.equal(case["pad"][1, g].cpu(), padded - (padded.view(torch.int32) & ~0x1FFF).view(torch.float32)), "its lo words"
    _check_fused(case)


def test_critic_forward_fused_two_jobs(rt):
    """Two passes in one launch: the online critics on the long batch (activations stored) beside the target critics on a
    short one (head only), as the CQL step launches them (policy/model_free/cql.py:108-160)."""
    a = _fused_case(rt, 7936, 256, 23, 3, 2, seed=1)
    b = _fused_case(rt, 256, 256, 23, 3, 2, seed=2, store_h=False)
    for pairs in (False, True):
        for t in (a["out"], b["out"]):
            t.fill_(float("nan"))
        rt.critic_fwd_fused([a["job"], b["job"]], pairs=pairs)()
        torch.cuda.synchronize()
        _check_fused(a)
        _check_fused(b)


@pytest.mark.parametrize("pairs", [False, True])
@pytest.mark.parametrize("M,N,K0,nh,G", [(7936, 256, 23, 3, 2), (300, 128, 32, 4, 1), (2048 + 77, 64, 5, 2, 3)])
def test_critic_backward_fused_matches_fp64(rt, M, N, K0, nh, G, pairs):
    """orlk_critic_bwd_fused behind orlk_critic_fwd_fused: dZ_l of every hidden layer but the last against fp64 autograd
    algebra on the activations the forward kernel stored (so that both sides take the same ReLU decisions)."""
    case = _fused_case(rt, M, N, K0, nh, G, seed=7 * M + nh)
    bits = torch.full((nh, G, 8, M), -1, dtype=torch.int32, device=DEV)
    case["job"]["relu_bits"] = bits.data_ptr()
    rt.critic_fwd_fused([case["job"]])()
    torch.cuda.synchronize()
    H = [h.cpu() for h in case["H"]]
    for l in range(nh):         # the decision bits are exactly (H > 0)
        want = (H[l] > 0).view(G, M, N // 32, 32).to(torch.int64)
        words = (want << torch.arange(32)).sum(-1)
        got = bits[l].cpu().to(torch.int64).transpose(1, 2) & 0xFFFFFFFF          # [G][M][8]
        assert torch.equal(got[..., :N // 32], words), f"relu bits of layer {l}"
    P, offs, block = case["P"], case["offs"], case["block"]
    WT = torch.zeros_like(P)
    for g in range(G):
        for l in range(1, nh):
            wo = g * block + offs[l][0]
            WT[wo:wo + N * N] = P[wo:wo + N * N].view(N, N).t().reshape(-1)
    WTd = WT.to(DEV)
    WTlo = torch.empty_like(WTd)
    rt.fused_prep(WTd, WTlo)()
    gen = torch.Generator().manual_seed(5)
    dq = torch.randn(G, M, generator=gen)
    dqd = dq.to(DEV)
    dZ = [torch.full((G, M, N), float("nan"), device=DEV) for _ in range(nh - 1)]
    Pd = case["keep"][1]
    op = rt.critic_bwd_fused(dq=dqd.data_ptr(), dq_gs=M, head_w=Pd.data_ptr() + 4 * offs[nh][0], relu_bits=bits.data_ptr(),
                             WT=[0] + [WTd.data_ptr() + 4 * offs[l][0] for l in range(1, nh)],
                             WTlo=[0] + [WTlo.data_ptr() + 4 * offs[l][0] for l in range(1, nh)],
                             dZ=[t.data_ptr() for t in dZ], gs=block, dz_gs=M * N, M=M, N=N, G=G, pairs=pairs)
    for _ in range(2):
        op()
    torch.cuda.synchronize()
    for g in range(G):
        Ws = case["Ws"][g]
        d = dq[g].double()[:, None] * Ws[nh].double() * (H[nh - 1][g] > 0)
        tol = 0.0
        for l in range(nh - 1, 0, -1):
            W = Ws[l].double()                               # [out, in]
            tol += 3e-6 * (d.abs() @ W.abs()).max().item()
            d = (d @ W) * (H[l - 1][g] > 0)
            got = dZ[l - 1][g]
            assert not torch.isnan(got).any(), f"dZ[{l - 1}] member {g}: unwritten output"
            err = (got.double().cpu() - d).abs().max().item()
            assert err <= tol, f"dZ[{l - 1}] member {g}: max err {err:.3e} vs {tol:.3e}"


def test_fused_prep_multi_equals_single(rt):
    """orlk_fused_prep_multi (one launch for several arenas) writes exactly what orlk_fused_prep writes per arena."""
    gen = torch.Generator().manual_seed(11)
    G, N, K0 = 2, 256, 23
    arenas = [torch.randn(n, generator=gen).to(DEV) for n in (G * (N * K0 + 1000), 70_001 * 4, 4096)]
    lo_a = [torch.full_like(a, float("nan")) for a in arenas]
    lo_b = [torch.full_like(a, float("nan")) for a in arenas]
    pad_a = torch.full((2, G, N, 32), float("nan"), device=DEV)
    pad_b = torch.full((2, G, N, 32), float("nan"), device=DEV)
    gs = N * K0 + 1000
    rt.fused_prep(arenas[0], lo_a[0], W0=arenas[0].data_ptr(), gs=gs, N=N, K0=K0, G=G, w0pad=pad_a)()
    rt.fused_prep(arenas[1], lo_a[1])()
    rt.fused_prep(arenas[2], lo_a[2])()
    rt.fused_prep_multi([dict(src=arenas[0], dst_lo=lo_b[0], W0=arenas[0].data_ptr(), gs=gs, N=N, K0=K0, G=G, w0pad=pad_b),
                         dict(src=arenas[1], dst_lo=lo_b[1]), dict(src=arenas[2], dst_lo=lo_b[2])])()
    torch.cuda.synchronize()
    for a, b in zip(lo_a + [pad_a], lo_b + [pad_b]):
        assert not torch.isnan(b).any() and torch.equal(a, b)
