"""EDAC with the critic ensemble sharded over ranks (BASELINE.json configs[2]).

Every rank holds the actor, alpha, the batch and the noise (replicated: same seeds, same indices) and a contiguous slice
of the E critics (10 over 4 ranks = 3/3/2/2).  The reference's ensemble-wide reductions become three exchanges per step:

  X1  all-gather q_e(s, a~pi)            [E, B]        -> argmin over ALL critics in the actor loss   (edac.py:96-102)
  X2  all-gather dQ/da of the own slice  [E, B, A]     -> the actor's upstream gradient, summed over e (edac.py:100)
  X3  all-gather [ Q'_e(s', a') | dQ_e/da (g-chain) | Q_e(s, a_data) ]
                                                       -> min_e of the TD target (edac.py:124-131), S = sum_e ghat_e of the
                                                          diversity term (:136-149), the logged sum of TD losses (:133-134)

Each is ONE equal-block all-gather (blocks padded to the largest slice) followed by ``orlk_compact_blocks``; everything
between two exchanges is one captured CUDA graph (four per step).  The actor / alpha updates are computed redundantly
and bit-identically on every rank (same gathered inputs, same summation order); each rank's Adam only touches its own
critic slice.  The exchange itself is behind ``comm.all_gather(send, recv)``: NCCL (``torch.distributed``) between
processes, plain device copies when several ranks are emulated on one GPU (tests)."""
import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import _lib as L
from .core import GP, Mat, Plan
from .edac import EDACLearner, LS_DIV, LS_TD_SUM
from .learner import (emit_dact, emit_forward, emit_head_dgrad, emit_hidden_dgrad, emit_wgrad_adam, make_gradbuf, wgrad_layout)
from .nets import GradBuf, TC_MIN_ROWS, adam_descs, ens_n_tile, pick_cfg
from .sac_family import LS_ACTOR, LS_ALPHA, LS_ALPHA_LOSS


class MemberSlice(nn.Module):
    """EnsembleLinear-like holder (``weight`` [E_r, in, out], ``bias`` [E_r, 1, out]) of members [e0, e1) of a layer."""

    def __init__(self, lay: nn.Module, e0: int, e1: int):
        super().__init__()
        self.src, self.e0, self.e1 = [lay], e0, e1          # (a list: not registered as a sub-module)
        self.num_ensemble = e1 - e0
        self.weight = nn.Parameter(lay.weight.detach()[e0:e1].clone())
        self.bias = nn.Parameter(lay.bias.detach()[e0:e1].clone())

    @torch.no_grad()
    def write_back(self) -> None:
        lay = self.src[0]
        lay.weight.data[self.e0:self.e1].copy_(self.weight.data)
        lay.bias.data[self.e0:self.e1].copy_(self.bias.data)


class NcclComm:
    """Equal-block all-gather between processes (``torch.distributed``, NCCL over NVLink) on torch's current stream --
    the stream the step graphs are launched on, so the exchange is ordered between them without host synchronisation."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, send: torch.Tensor, recv: torch.Tensor) -> None:
        self.dist.all_gather_into_tensor(recv, send, group=self.group)


class EDACShardedLearner(EDACLearner):
    def __init__(self, policy, batch_size: int, rank: int, world: int, comm=None, seed: int = 0):
        from ..parallel import partition_members
        self.rank, self.world, self.comm = int(rank), int(world), comm
        self.E_all = int(policy.critics._num_ensemble)
        parts = partition_members(self.E_all, self.world)
        if min(len(p) for p in parts) == 0 or self.world > 8:
            raise L.OrlkError(f"cannot shard {self.E_all} critics over {self.world} ranks (1..8 ranks, at least one critic each)")
        self.counts = [len(p) for p in parts]
        self.e0, self.e1 = parts[self.rank][0], parts[self.rank][-1] + 1
        self.E_max = max(self.counts)
        self._slices: List[MemberSlice] = []
        super().__init__(policy, batch_size, seed)

    # EDACLearner.__init__ builds the critic ParamSet from these: the local members only
    def _critic_layers(self, policy):
        lin = lambda m: [x for x in m.model if hasattr(x, "num_ensemble")]
        cur = [MemberSlice(x, self.e0, self.e1) for x in lin(policy.critics)]
        old = [MemberSlice(x, self.e0, self.e1) for x in lin(policy.critics_old)]
        self._slices = cur + old
        return cur, old

    def write_back(self) -> None:
        """Copy the trained slice into the policy's full EnsembleCritic modules (other ranks' members are untouched)."""
        for s in self._slices:
            s.write_back()

    @torch.no_grad()
    def gather_all(self) -> None:
        """Make the policy's full ``critics`` / ``critics_old`` modules hold EVERY rank's trained members (checkpoints,
        ``state_dict()``): one padded all-gather per parameter tensor."""
        self.write_back()
        if self.comm is None or self.world == 1:
            return
        self.rt.sync()
        for s in self._slices:
            lay = s.src[0]
            for name in ("weight", "bias"):
                full = getattr(lay, name).data
                per = full[0].numel()
                send = torch.zeros(self.E_max * per, dtype=torch.float32, device=full.device)
                send[:self.E * per].copy_(getattr(s, name).data.reshape(-1))
                recv = torch.empty(self.world * self.E_max * per, dtype=torch.float32, device=full.device)
                self.comm.all_gather(send, recv)
                e = 0
                for r, c in enumerate(self.counts):
                    blk = recv[r * self.E_max * per:r * self.E_max * per + c * per]
                    full[e:e + c].copy_(blk.view((c,) + tuple(full.shape[1:])))
                    e += c

    # ------------------------------------------------------------------ exchange buffers
    def _xbuf(self, per_member: Sequence[int]):
        """(send [E_max * sum(per_member)], recv [world, same]) for one exchange of several per-member arrays."""
        n = self.E_max * sum(per_member)
        return self.rt.zeros(n), self.rt.zeros(self.world * n)

    def _compact(self, plan: Plan, tag: str, recv: torch.Tensor, block: int, off: int, per_member: int, dst: torch.Tensor) -> None:
        cnt = (C.c_int * 8)(*(self.counts + [0] * (8 - self.world)))
        args = (recv.data_ptr() + 4 * off, block, dst.data_ptr(), self.world, per_member, cnt)
        plan.keep.append(cnt)
        plan.add(tag, lambda: L.call("orlk_compact_blocks", *args, self.rt.cur))

    def _stage_local(self, plan: Plan, tag: str, src: torch.Tensor, send: torch.Tensor, off: int, n: int) -> None:
        """local slice (n floats) -> its place in the send block"""
        args = (send.data_ptr() + 4 * off, src.data_ptr(), 4 * n)
        plan.add(tag, lambda: L.call("orlk_memcpy_d2d_async", *args, self.rt.cur))

    # ------------------------------------------------------------------ the four graph segments
    def _build(self) -> None:
        rt, B, O, A, pol = self.rt, self.B, self.O, self.A, self.policy
        E, Ea, Em = self.E, self.E_all, self.E_max          # local, global, largest slice
        e0, e1 = self.e0, self.e1
        cps, aps = self.critic_ps, self.actor_ps
        nh = self.nh_c
        Bt = B * self.n_next
        run_a = self.mlp_run(aps, B, self.nh_a, need_grad=True)
        run_ca = self.mlp_run(cps, B, nh, need_grad=True)
        run_an = self.mlp_run(aps, B, self.nh_a, need_grad=False)
        run_t = self.mlp_run(cps, Bt, nh, need_grad=False, store="T")
        run_c = self.mlp_run(cps, B, nh, need_grad=True)
        run_g = self.mlp_run(cps, B, nh, need_grad=True, share_forward=run_c)
        ldx = (O + A + 3) // 4 * 4
        Xa, Xt, Xd = rt.zeros(B, ldx)[:, :O + A], rt.zeros(Bt, ldx)[:, :O + A], rt.zeros(B, ldx)[:, :O + A]
        logp_a, lp_next, glp = rt.zeros(B), rt.zeros(Bt), rt.zeros(B)
        tq_best = rt.zeros(E, B) if self.n_next > 1 else None
        dA = rt.zeros(E, B, A)
        gin = rt.zeros(E, B, A)
        ones = torch.ones(B, 1, dtype=torch.float32, device=self.dev)
        ubar = [rt.zeros(E, B, cps.layers[l].out_dim) for l in range(nh)]
        div_scratch = rt.zeros((B + 255) // 256)
        # ensemble-wide (dense, all E_all members) views of what the exchanges deliver
        q_all, dq_all = rt.zeros(Ea, B), rt.zeros(Ea, B)
        dA_all = rt.zeros(Ea, B, A)
        tq_all, gin_all, gbar_all = rt.zeros(Ea, B), rt.zeros(Ea, B, A), rt.zeros(Ea, B, A)
        qd_all, dqd_all = rt.zeros(Ea, B), rt.zeros(Ea, B)
        # the local members' upstream gradients ARE slices of the ensemble-wide results
        run_ca.dOut = dq_all[e0:e1].view(E, B, 1)
        run_c.dOut = dqd_all[e0:e1].view(E, B, 1)
        gbar = gbar_all[e0:e1]
        s1, r1 = self._xbuf([B])
        s2, r2 = self._xbuf([B * A])
        s3, r3 = self._xbuf([B, B * A, B])
        blk1, blk2, blk3 = Em * B, Em * B * A, Em * (B + B * A + B)
        gb_a = make_gradbuf(rt, aps, [run_a])
        td_layout = wgrad_layout(cps, nh + 1, B)
        s_td = max(s for _, s in td_layout)
        gb_c = GradBuf(rt, cps, s_td + 1)
        obs2 = Mat.of(self.obs2)
        obs, nobs = obs2.rows_(0, B), obs2.rows_(B, 2 * B)
        mXa, mXt, mXd = Mat.of(Xa), Mat.of(Xt), Mat.of(Xd)

        # ---- S1: a ~ pi(s), the own critics at (s, a)                                          (edac.py:96-97)
        p1 = Plan(rt, "edac.s1")
        nargs = (self.noise.data_ptr(), (1 + self.n_next) * B * A, 0, 0.0, 1.0, int(self.seed), self.philox_counter.data_ptr(),
                 self.noise_enable.data_ptr())
        p1.add("philox", lambda: L.call("orlk_philox_fill", *nargs, rt.cur))
        emit_forward(rt, p1, run_a, [obs], "A.actor")
        self._sample(p1, "A.sample", run_a.out[0], self.noise_views["eps_actor"], mXa, logp_a, obs)
        emit_forward(rt, p1, run_ca, [mXa] * E, "A.critics")
        self._stage_local(p1, "X1.stage", run_ca.out, s1, 0, E * B)

        # ---- S2: actor loss over ALL critics, backward through the own ones                    (edac.py:98-102)
        p2 = Plan(rt, "edac.s2")
        self._compact(p2, "X1.compact", r1, blk1, 0, B, q_all)
        largs = (q_all.data_ptr(), B, Ea, logp_a.data_ptr(), B, self.scalars.data_ptr(), int(self.auto_alpha), 1,
                 self.target_entropy, self.groups_ptr, max(self.g_alpha, 0), self.alpha_mv.data_ptr(), dq_all.data_ptr(), B,
                 glp.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_ACTOR)
        p2.add("A.loss", lambda: L.call("orlk_sac_actor_loss", *largs, rt.cur))
        emit_head_dgrad(rt, p2, run_ca, "A.critics")
        emit_hidden_dgrad(rt, p2, run_ca, "A.critics")
        emit_dact(rt, p2, run_ca, dA, O, A, "A.critics")
        self._stage_local(p2, "X2.stage", dA, s2, 0, E * B * A)

        # ---- S3: actor update (replicated); then a' ~ pi_new(s'), own target critics; own critics on the data rows and the
        #          first half of the diversity term's chain on a parallel branch                (edac.py:100-102, 112-147)
        p3 = Plan(rt, "edac.s3")
        self._compact(p3, "X2.compact", r2, blk2, 0, B * A, dA_all)
        bargs = (run_a.out.data_ptr(), 2 * A, self.noise_views["eps_actor"].data_ptr(), mXa.ptr + 4 * O, mXa.ld, dA_all.data_ptr(), Ea,
                 B * A, A, glp.data_ptr(), B, A, run_a.dOut.data_ptr(), 2 * A)
        p3.add("A.head_bwd", lambda: L.call("orlk_tanh_gauss_bwd", *bargs, rt.cur))
        par = B < TC_MIN_ROWS and os.environ.get("ORLK_EDAC_BRANCHES", "1") != "0"

        def data_and_g():
            p3.add("C.concat", rt.concat([(mXd, obs, 1, Mat.of(self.act))]))
            emit_forward(rt, p3, run_c, [mXd] * E, "C.critics")
            self._stage_local(p3, "X3.stage_qd", run_c.out, s3, Em * (B + B * A), E * B)
            if self.eta > 0:
                run_g.dOut.fill_(1.0)
                emit_head_dgrad(rt, p3, run_g, "G.chain")
                emit_hidden_dgrad(rt, p3, run_g, "G.chain")
                emit_dact(rt, p3, run_g, gin, O, A, "G")
                self._stage_local(p3, "X3.stage_gin", gin, s3, Em * B, E * B * A)

        if par:
            p3.fork()
            p3.branch(1)
            data_and_g()
            p3.branch(0)
        emit_head_dgrad(rt, p3, run_a, "A.actor")
        emit_hidden_dgrad(rt, p3, run_a, "A.actor")
        emit_wgrad_adam(rt, p3, run_a, [obs], gb_a, self.groups_ptr, "A.actor", polyak=False)
        emit_forward(rt, p3, run_an, [nobs], "C.actor_next")
        self._sample(p3, "C.sample_next", run_an.out[0], self.noise_views["eps_next"], mXt, lp_next, nobs, rep=self.n_next)
        emit_forward(rt, p3, run_t, [mXt] * E, "C.target")
        tq = run_t.out
        if self.n_next > 1:
            margs = (run_t.out.data_ptr(), Bt, E, B, self.n_next, tq_best.data_ptr(), B)
            p3.add("C.target_max", lambda: L.call("orlk_segment_max", *margs, rt.cur))
            tq = tq_best
        self._stage_local(p3, "X3.stage_tq", tq, s3, 0, E * B)
        if par:
            p3.join()
        else:
            data_and_g()

        # ---- S4: TD target from ALL target critics, diversity term over ALL critics, update of the own slice
        p4 = Plan(rt, "edac.s4")
        self._compact(p4, "X3.compact_tq", r3, blk3, 0, B, tq_all)
        self._compact(p4, "X3.compact_qd", r3, blk3, Em * (B + B * A), B, qd_all)
        div_terms = None
        if self.eta > 0:
            self._compact(p4, "X3.compact_gin", r3, blk3, Em * B, B * A, gin_all)
            dargs = (gin_all.data_ptr(), Ea, B, A, self.eta, gbar_all.data_ptr(), div_scratch.data_ptr(),
                     self.loss_dev.data_ptr() + 4 * LS_DIV)
            p4.add("G.div", lambda: L.call("orlk_edac_div", *dargs, rt.cur))
            div_terms = [(0, gbar, run_g.dZ[0])]
            for l in range(nh):
                lay = cps.layers[l]
                if l >= 1 and run_c.ens_tc and run_c.tc_fwd[l]:
                    K, N = lay.in_dim, lay.out_dim
                    p4.add(f"G.ubar{l}.tc", rt.tc_gemm(
                        A=Mat(ubar[l - 1].data_ptr(), B, K, K), a_gs=B * K, B=Mat(cps.w(l, 0), K, N, N), b_gs=lay.w_gs,
                        b_mn=True, G=E, passes=run_c.tc, n_tile=ens_n_tile(E, B, N), epi=L.EPI_RELU_MASK,
                        C=Mat(ubar[l].data_ptr(), B, N, N), c_gs=B * N, aux=Mat(run_c.H[l].data_ptr(), B, N, N), aux_gs=B * N))
                else:
                    fprobs = []
                    for e in range(E):
                        if l == 0:
                            fprobs.append(GP(A=gbar[e].data_ptr(), lda=A, a_layout=0, B=cps.w(0, e) + 4 * O * lay.out_dim,
                                             ldb=lay.out_dim, b_layout=0, C=ubar[0][e].data_ptr(), ldc=lay.out_dim, M=B,
                                             N=lay.out_dim, K=A, epi=L.EPI_RELU_MASK, aux=run_c.H[0][e].data_ptr(),
                                             ldaux=lay.out_dim))
                        else:
                            fprobs.append(GP(A=ubar[l - 1][e].data_ptr(), lda=lay.in_dim, a_layout=0, B=cps.w(l, e),
                                             ldb=lay.out_dim, b_layout=0, C=ubar[l][e].data_ptr(), ldc=lay.out_dim, M=B,
                                             N=lay.out_dim, K=lay.in_dim, epi=L.EPI_RELU_MASK, aux=run_c.H[l][e].data_ptr(),
                                             ldaux=lay.out_dim))
                    p4.add(f"G.ubar{l}", rt.gemm(fprobs, pick_cfg(B * E, lay.out_dim, rows_per_problem=B)))
                div_terms.append((l + 1, ubar[l], run_g.dZ[l + 1] if l + 1 < nh else None))
        use_alpha = 0 if (pol._deterministic_backup or self.n_next > 1) else 1
        # over ALL members: every rank logs the same sum of TD losses; the own members' dq is a slice of dqd_all
        targs = (qd_all.data_ptr(), B, Ea, tq_all.data_ptr(), B, Ea, lp_next.data_ptr(), self.scalars.data_ptr(), use_alpha,
                 self.rew.data_ptr(), self.term.data_ptr(), B, float(pol._gamma), dqd_all.data_ptr(), B, None,
                 self.td_each.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_TD_SUM)
        p4.add("C.td_loss", lambda: L.call("orlk_td_loss", *targs, rt.cur))
        emit_head_dgrad(rt, p4, run_c, "C.critics")
        emit_hidden_dgrad(rt, p4, run_c, "C.critics")
        td_terms = [(l, (mXd if l == 0 else run_c.H[l - 1]), (run_c.dZ[l] if l < nh else run_c.dOut)) for l in range(nh + 1)]
        self._emit_ens_wgrads(p4, "C.critics.wgrad_td", run_c, gb_c, td_terms, ones, split_base=0, with_bias=True,
                              k_splits=td_layout, tiny=(B < TC_MIN_ROWS and s_td == 1))
        if self.eta > 0:
            self._emit_ens_wgrads(p4, "G.wgrad_div", run_c, gb_c, div_terms, ones, split_base=s_td, with_bias=False,
                                  k_splits=None, tiny=B < TC_MIN_ROWS, col0_first=O)
        splits = [s_td + 1] * (nh + 1)
        p4.add("C.critics.adam", rt.adam(adam_descs(cps, gb_c, splits, polyak=True), self.groups_ptr))
        mask = (1 << self.g_actor) | (1 << self.g_c) | ((1 << self.g_alpha) if self.g_alpha >= 0 else 0)
        self.finish_ops(p4, mask)
        p4.keep.append(dict(locals()))
        self.plans.update({"s1": p1, "s2": p2, "s3": p3, "s4": p4})
        self.segments = ("s1", "s2", "s3", "s4")
        self.exchanges = ((s1, r1), (s2, r2), (s3, r3))
        self._built = True

    def _rebatch(self, B: int) -> None:
        super()._rebatch(B)
        self.td_each = self.rt.zeros(self.E_all)

    # ------------------------------------------------------------------ stepping
    def prepare(self, batch, noise=None) -> None:
        self.bind_batch(batch)
        if not self._built:
            self._build()
        self.set_noise(noise)
        self.sync_lr()
        self.refresh()
        tok = getattr(self, "_bound_token", None)
        if tok is not None and getattr(tok, "pending", False):
            tok.materialise()

    def run_segment(self, i: int) -> None:
        plan = self.plans[self.segments[i]]
        if self.use_graph:
            plan.launch()
        else:
            plan.run_eager()

    def finish(self) -> Dict[str, float]:
        self.rt.sync()
        self.steps_done += 1
        out = self.loss_np.tolist()
        res = {"loss/actor": float(out[LS_ACTOR]),
               "loss/critics": float(out[LS_TD_SUM]) + (float(out[LS_DIV]) if self.eta > 0 else 0.0)}
        if self.auto_alpha:
            res["loss/alpha"] = float(out[LS_ALPHA_LOSS])
            res["alpha"] = float(out[LS_ALPHA])
        return res

    def learn_many(self, buffer, n_steps: int) -> List[Dict[str, float]]:
        """(the sharded step is several graphs with collectives in between: K plain steps)"""
        return [self.step(buffer.sample(self.B)) for _ in range(int(n_steps))]

    def _capture_whole_step(self) -> bool:
        """The four segments AND the three NCCL all-gathers as ONE CUDA graph (torch's stream capture records the
        collectives' kernels like any other launch): one graph launch per step instead of four launches and three Python
        collective calls.  ORLK_SHARD_GRAPH=0 keeps the segmented form."""
        if os.environ.get("ORLK_SHARD_GRAPH", "1") == "0" or not self.use_graph:
            return False
        rt = self.rt
        for i in range(len(self.segments)):             # warm-up outside capture: NCCL sets its channels up lazily
            self.plans[self.segments[i]].run_eager()
            if i < len(self.exchanges):
                self.comm.all_gather(*self.exchanges[i])
        rt.sync()
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                rt.cur = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
                for i in range(len(self.segments)):
                    self.plans[self.segments[i]].run_eager()
                    if i < len(self.exchanges):
                        self.comm.all_gather(*self.exchanges[i])
        finally:
            rt.cur = rt.exec_ptr
        self._whole = g
        return True

    def step(self, batch, noise=None) -> Dict[str, float]:
        if self.comm is None:
            raise L.OrlkError("EDACShardedLearner.step needs a communicator (NcclComm), or drive the ranks with EmulatedShardGroup")
        self.prepare(batch, noise)
        whole = getattr(self, "_whole", None)
        if whole is None and not getattr(self, "_whole_tried", False):
            # (the warm-up pass inside _capture_whole_step is a real step: it advances the optimiser like any other)
            self._whole_tried = True
            if self._capture_whole_step():
                return self.finish()
            whole = None
        if whole is not None:
            whole.replay()
            return self.finish()
        for i in range(len(self.segments)):
            self.run_segment(i)
            if i < len(self.exchanges):
                send, recv = self.exchanges[i]
                self.comm.all_gather(send, recv)
        return self.finish()


class EmulatedShardGroup:
    """All ranks of a member-sharded run inside ONE process on ONE GPU (tests, and G > number of GPUs): the rank engines run
    segment by segment in lockstep and an exchange is ``world`` device copies per rank.  (Separate processes that spin on
    each other's kernels must not share a GPU; here nothing waits: the segments are ordinary stream-ordered launches.)"""

    def __init__(self, engines: Sequence[EDACShardedLearner]):
        self.engines = list(engines)
        assert [e.rank for e in self.engines] == list(range(len(self.engines)))

    def step(self, batches, noise=None) -> List[Dict[str, float]]:
        engs = self.engines
        for e, b in zip(engs, batches):
            e.prepare(b, noise)
        for i in range(len(engs[0].segments)):
            for e in engs:
                e.run_segment(i)
            if i < len(engs[0].exchanges):
                for e in engs:
                    _, recv = e.exchanges[i]
                    n = recv.numel() // len(engs)
                    for q in engs:
                        recv[q.rank * n:(q.rank + 1) * n].copy_(q.exchanges[i][0])
        return [e.finish() for e in engs]
