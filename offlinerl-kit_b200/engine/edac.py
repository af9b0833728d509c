"""Gradient-step engine for EDAC (policy/model_free/edac.py:88-166).

E critics live in one 'io' ParamSet (EnsembleLinear layout, members contiguous).  The ensemble-diversity term needs
d/dW of a function of the INPUT gradients dQ_e/da; instead of autograd's double backward (``create_graph=True``) the
closed form of SURVEY.md appendix A.4 is scheduled explicitly:
  1. g-chain   : an ordinary backward pass with upstream 1 over the saved ReLU masks -> v_l and g = dQ/da
  2. orlk_edac_div : loss eta*G and gbar = eta * dG/dg
  3. second chain  : ubar_1 = m_1 * (gbar W1a), ubar_{l+1} = m_{l+1} * (ubar_l W_{l+1})  (forward-like masked GEMMs)
                     dW1a += gbar^T v_1, dW_{l+1} += ubar_l^T v_{l+1}, dw_head += sum_b ubar_L      (wgrad GEMMs)
The diversity weight gradients go to extra split-K slots of the same GradBuf, so one fused Adam+polyak launch sums
the TD part and the diversity part.
"""
import ctypes as C
import os
from typing import Dict

import torch

from .. import _lib as L
from .core import GP, Mat, Plan
from .learner import (Learner, MlpRun, check_plain_mlp, emit_dact, emit_forward, emit_head_dgrad, emit_hidden_dgrad,
                      emit_wgrad_adam, linears_of, make_gradbuf, wgrad_layout)
from .nets import GradBuf, ParamSet, TC_MIN_ROWS, adam_descs, dgrad_problem, ens_n_tile, pick_cfg, wgrad_problem
from .sac_family import LS_ACTOR, LS_ALPHA, LS_ALPHA_LOSS
from .td3_iql import _BatchMixin

LS_TD_SUM, LS_DIV = 9, 10


class EDACLearner(_BatchMixin, Learner):
    def __init__(self, policy, batch_size: int, seed: int = 0):
        actor = policy.actor
        super().__init__(actor.device)
        rt = self.rt
        self.policy, self.B, self.seed = policy, int(batch_size), seed
        check_plain_mlp(actor.backbone, "actor")
        # max_q_backup (edac.py:113-122): 10 sampled next actions per row (the count is a literal there), each target
        # critic maximised over them, no entropy term
        self.n_next = 10 if policy._max_q_backup else 1
        dist = actor.dist_net
        if (getattr(dist, "_sigma_min", -5.0), getattr(dist, "_sigma_max", 2.0)) != (-5.0, 2.0):
            # the sampler / backward kernels clamp log sigma to the reference's default [-5, 2] (dist_module.py:95-105)
            raise L.OrlkError("the CUDA engine implements TanhDiagGaussian with sigma_min=-5, sigma_max=2 only")
        if not (getattr(dist, "_c_sigma", False) and getattr(dist, "_unbounded", False)):
            raise L.OrlkError("EDAC engine needs TanhDiagGaussian(unbounded=True, conditioned_sigma=True)")
        for mod in policy.critics.model:
            if not hasattr(mod, "num_ensemble") and not isinstance(mod, torch.nn.ReLU):
                raise L.OrlkError(f"EnsembleCritic: only ReLU activations are supported, found {type(mod).__name__}")
        self.actor_ps = ParamSet.from_linear_members(rt, "actor", [linears_of(actor)], fuse_last=2)
        cur, old = self._critic_layers(policy)
        self.critic_ps = ParamSet.from_ensemble(rt, "critics", cur, targets=old)
        self.param_sets = [self.actor_ps, self.critic_ps]
        self.E = self.critic_ps.G
        self.nh_a, self.nh_c = len(self.actor_ps.layers) - 1, len(self.critic_ps.layers) - 1
        self.O = self.actor_ps.layers[0].in_dim
        self.A = self.actor_ps.layers[-1].out_dim // 2
        self.g_actor = self.add_group(policy.actor_optim)
        self.g_c = self.add_group(policy.critics_optim, tau=float(policy._tau))
        self.actor_ps.group_ids, self.critic_ps.group_ids = [self.g_actor], [self.g_c]
        self.auto_alpha = bool(policy._is_auto_alpha)
        self.alpha_mv = rt.zeros(2)
        self.g_alpha, self.target_entropy = -1, 0.0
        with torch.no_grad():
            self.scalars[L.SC_ALPHA] = float(policy._alpha)
        if self.auto_alpha:
            self.g_alpha = self.add_group(policy.alpha_optim)
            self.target_entropy = float(policy._target_entropy)
            la = policy._log_alpha
            with torch.no_grad():
                self.scalars[L.SC_LOG_ALPHA] = float(la.detach().reshape(-1)[0])
            la.data = self.scalars[L.SC_LOG_ALPHA:L.SC_LOG_ALPHA + 1].view(la.shape)
        self.eta = float(policy._eta)
        self.push_groups()
        self._rebatch(self.B)

    def _rebatch(self, B: int) -> None:
        rt, A = self.rt, self.A
        self.B = B
        self._make_stage()
        self.noise = rt.zeros((1 + self.n_next) * B * A)
        self.noise_views = {"eps_actor": self.noise[:B * A].view(B, A),
                            "eps_next": self.noise[B * A:].view(self.n_next * B, A)}
        self._built = False

    def _critic_layers(self, policy):
        """(EnsembleLinear layers of the online critics, of the target critics) this engine trains: all members here,
        a contiguous slice of them in the member-sharded engine (engine/edac_sharded.py)."""
        lin = lambda m: [x for x in m.model if hasattr(x, "num_ensemble")]
        return lin(policy.critics), lin(policy.critics_old)

    def _sample(self, plan, tag, head, eps, X: Mat, logp, obs: Mat, rep: int = 1):
        O, A, B = self.O, self.A, self.B
        args = (head.data_ptr(), 2 * A, 0, rep, eps.data_ptr(), B * rep, A, X.ptr + 4 * O, X.ld, logp.data_ptr(), obs.ptr, obs.ld,
                O, X.ptr, X.ld)
        plan.add(tag, lambda: L.call("orlk_tanh_gauss_sample", *args, self.rt.cur))

    def _emit_ens_wgrads(self, plan, tag, run, gb, terms, ones, split_base, with_bias, k_splits, tiny, col0_first=0):
        """Weight (and bias) gradients of the critic ensemble for a list of (layer, X, dY) terms, X [E,B,in] (or a Mat
        shared by the members), dY [E,B,out] (None: a column of ones, i.e. column sums of X).  Square hidden layers go
        to the tensor-core kernel, one launch per layer over all members with both operands read MN-major as they are
        stored (dW[i][o] = sum_m X[m][i] dY[m][o]); their bias gradients and the narrow layers stay on the grouped
        small-row kernel; the launches are independent and run on parallel branches of the step graph."""
        rt, cps, E, B = self.rt, self.critic_ps, self.E, self.B
        tc_launches, probs = [], []
        for (l, X, dY) in terms:
            lay = cps.layers[l]
            shared = isinstance(X, Mat)
            on_tc = (run.ens_tc and not shared and dY is not None and 1 <= l < run.nh and lay.in_dim % 32 == 0
                     and lay.out_dim % 32 == 0 and lay.in_dim <= 256 and ens_n_tile(E, lay.in_dim, lay.out_dim) > 0
                     and os.environ.get("ORLK_ENSEMBLE_TC_WGRAD", "1") != "0")
            if on_tc:
                i, o = lay.in_dim, lay.out_dim
                tc_launches.append((f"{tag}{l}.tc", rt.tc_gemm(
                    A=Mat(X.data_ptr(), B, i, i), a_gs=B * i, a_mn=True, B=Mat(dY.data_ptr(), B, o, o), b_gs=B * o, b_mn=True,
                    G=E, passes=run.tc, n_tile=ens_n_tile(E, i, o),
                    C=Mat(gb.ptr(lay.w_off) + 4 * split_base * gb.stride, i, o, o), c_gs=lay.w_gs)))
                if with_bias:       # db[o] = sum_m dY[m][o]: a one-row product with the column of ones
                    for e in range(E):
                        probs.append(GP(A=ones.data_ptr(), lda=1, a_layout=1, B=dY[e].data_ptr(), ldb=o, b_layout=0,
                                        C=gb.ptr(lay.b_off + e * lay.b_gs), ldc=o, M=1, N=o, K=B, k_splits=1,
                                        split_base=split_base, c_split_stride=gb.stride, sum_split_stride=gb.stride))
                continue
            for e in range(E):
                x = X if shared else Mat.of(X[e])
                dy = Mat.of(ones) if dY is None else Mat.of(dY[e])
                ks = k_splits[l][1] if k_splits is not None else 1
                probs.append(wgrad_problem(cps, gb, l, e, x, dy, ks, split_base=split_base,
                                           col0=col0_first if (l == 0 and col0_first) else 0, with_bias=with_bias))
        rest = rt.gemm(probs, L.CFG_TINY if tiny else L.CFG_SMALL)
        if not tc_launches:
            plan.add(tag, rest)
            return
        plan.fork()
        for b, (label, op) in enumerate(tc_launches, start=1):
            plan.branch(b)
            plan.add(label, op)
        plan.branch(0)
        plan.add(tag, rest)
        plan.join()

    def _build(self) -> None:
        rt, B, O, A, E, pol = self.rt, self.B, self.O, self.A, self.E, self.policy
        cps, aps = self.critic_ps, self.actor_ps
        nh = self.nh_c
        run_a = self.mlp_run(aps, B, self.nh_a, need_grad=True)
        run_ca = self.mlp_run(cps, B, nh, need_grad=True)                       # Q_e(s, a~pi) in the actor phase
        run_an = self.mlp_run(aps, B, self.nh_a, need_grad=False)
        Bt = B * self.n_next
        run_t = self.mlp_run(cps, Bt, nh, need_grad=False, store="T")
        run_c = self.mlp_run(cps, B, nh, need_grad=True)                        # Q_e(s, a_data): TD backward
        run_g = self.mlp_run(cps, B, nh, need_grad=True, share_forward=run_c)   # same activations: input-gradient chain
        ldx = (O + A + 3) // 4 * 4          # 16-byte rows: the critics' first layer reads them through TMA
        Xa, Xt, Xd = rt.zeros(B, ldx)[:, :O + A], rt.zeros(Bt, ldx)[:, :O + A], rt.zeros(B, ldx)[:, :O + A]
        logp_a, lp_next, glp = rt.zeros(B), rt.zeros(Bt), rt.zeros(B)
        tq_best = rt.zeros(E, B) if self.n_next > 1 else None
        dA = rt.zeros(E, B, A)
        gin, gbar = rt.zeros(E, B, A), rt.zeros(E, B, A)
        ones = torch.ones(B, 1, dtype=torch.float32, device=self.dev)
        ubar = [rt.zeros(E, B, cps.layers[l].out_dim) for l in range(nh)]
        div_scratch = rt.zeros((B + 255) // 256)
        self._keep = [run_a, run_ca, run_an, run_t, run_c, run_g, Xa, Xt, Xd, logp_a, lp_next, glp, dA, gin, gbar, ones, ubar,
                      div_scratch, tq_best]
        gb_a = make_gradbuf(rt, aps, [run_a])
        td_layout = wgrad_layout(cps, nh + 1, B)
        s_td = max(s for _, s in td_layout)
        gb_c = GradBuf(rt, cps, s_td + 1)              # slots [0, s_td): TD part; slot s_td: diversity part
        obs2 = Mat.of(self.obs2)
        obs, nobs = obs2.rows_(0, B), obs2.rows_(B, 2 * B)
        mXa, mXt, mXd = Mat.of(Xa), Mat.of(Xt), Mat.of(Xd)
        plan = Plan(rt, "edac")
        nargs = (self.noise.data_ptr(), (1 + self.n_next) * B * A, 0, 0.0, 1.0, int(self.seed), self.philox_counter.data_ptr(),
                 self.noise_enable.data_ptr())
        plan.add("philox", lambda: L.call("orlk_philox_fill", *nargs, rt.cur))

        def actor_phase():
            # ---- actor: L = -mean min_e Q_e(s, a) + alpha mean logp      (edac.py:96-110)
            emit_forward(rt, plan, run_a, [obs], "A.actor")
            self._sample(plan, "A.sample", run_a.out[0], self.noise_views["eps_actor"], mXa, logp_a, obs)
            emit_forward(rt, plan, run_ca, [mXa] * E, "A.critics")
            largs = (run_ca.out.data_ptr(), B, E, logp_a.data_ptr(), B, self.scalars.data_ptr(), int(self.auto_alpha), 1,
                     self.target_entropy, self.groups_ptr, max(self.g_alpha, 0), self.alpha_mv.data_ptr(), run_ca.dOut.data_ptr(), B,
                     glp.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_ACTOR)
            plan.add("A.loss", lambda: L.call("orlk_sac_actor_loss", *largs, rt.cur))
            emit_head_dgrad(rt, plan, run_ca, "A.critics")
            emit_hidden_dgrad(rt, plan, run_ca, "A.critics")
            emit_dact(rt, plan, run_ca, dA, O, A, "A.critics")
            bargs = (run_a.out.data_ptr(), 2 * A, self.noise_views["eps_actor"].data_ptr(), mXa.ptr + 4 * O, mXa.ld, dA.data_ptr(), E,
                     B * A, A, glp.data_ptr(), B, A, run_a.dOut.data_ptr(), 2 * A)
            plan.add("A.head_bwd", lambda: L.call("orlk_tanh_gauss_bwd", *bargs, rt.cur))
            emit_head_dgrad(rt, plan, run_a, "A.actor")
            emit_hidden_dgrad(rt, plan, run_a, "A.actor")
            emit_wgrad_adam(rt, plan, run_a, [obs], gb_a, self.groups_ptr, "A.actor", polyak=False)


        def target_phase():
            # ---- critics: TD to min_e Q'_e(s', a') - alpha logp'  +  eta * diversity      (edac.py:112-155)
            emit_forward(rt, plan, run_an, [nobs], "C.actor_next")
            self._sample(plan, "C.sample_next", run_an.out[0], self.noise_views["eps_next"], mXt, lp_next, nobs, rep=self.n_next)
            emit_forward(rt, plan, run_t, [mXt] * E, "C.target")
            tq = run_t.out
            if self.n_next > 1:
                margs = (run_t.out.data_ptr(), Bt, E, B, self.n_next, tq_best.data_ptr(), B)
                plan.add("C.target_max", lambda: L.call("orlk_segment_max", *margs, rt.cur))
                tq = tq_best
            return tq

        def data_forward():
            plan.add("C.concat", rt.concat([(mXd, obs, 1, Mat.of(self.act))]))
            emit_forward(rt, plan, run_c, [mXd] * E, "C.critics")

        def g_chain():
            # 1. input-gradient chain (upstream 1)
            run_g.dOut.fill_(1.0)
            emit_head_dgrad(rt, plan, run_g, "G.chain")
            emit_hidden_dgrad(rt, plan, run_g, "G.chain")
            emit_dact(rt, plan, run_g, gin, O, A, "G")
            # 2. loss and gbar
            dargs = (gin.data_ptr(), E, B, A, self.eta, gbar.data_ptr(), div_scratch.data_ptr(),
                     self.loss_dev.data_ptr() + 4 * LS_DIV)
            plan.add("G.div", lambda: L.call("orlk_edac_div", *dargs, rt.cur))
            # 3. second chain + weight gradients into slot s_td
            div_terms = [(0, gbar, run_g.dZ[0])]
            for l in range(nh):
                lay = cps.layers[l]
                if l >= 1 and run_c.ens_tc and run_c.tc_fwd[l]:
                    # ubar_{l+1} = m_{l+1} * (ubar_l W_{l+1}) for all members in one tensor-core launch
                    K, N = lay.in_dim, lay.out_dim
                    plan.add(f"G.ubar{l}.tc", rt.tc_gemm(
                        A=Mat(ubar[l - 1].data_ptr(), B, K, K), a_gs=B * K, B=Mat(cps.w(l, 0), K, N, N), b_gs=lay.w_gs,
                        b_mn=True, G=E, passes=run_c.tc, n_tile=ens_n_tile(E, B, N), epi=L.EPI_RELU_MASK,
                        C=Mat(ubar[l].data_ptr(), B, N, N), c_gs=B * N, aux=Mat(run_c.H[l].data_ptr(), B, N, N), aux_gs=B * N))
                else:
                    fprobs = []
                    for e in range(E):
                        if l == 0:      # ubar_1 = m_1 * (gbar W1a),  W1a = rows O.. of W1 [in, out]
                            fprobs.append(GP(A=gbar[e].data_ptr(), lda=A, a_layout=0, B=cps.w(0, e) + 4 * O * lay.out_dim,
                                             ldb=lay.out_dim, b_layout=0, C=ubar[0][e].data_ptr(), ldc=lay.out_dim, M=B,
                                             N=lay.out_dim, K=A, epi=L.EPI_RELU_MASK, aux=run_c.H[0][e].data_ptr(),
                                             ldaux=lay.out_dim))
                        else:           # ubar_{l+1} = m_{l+1} * (ubar_l W_{l+1})
                            fprobs.append(GP(A=ubar[l - 1][e].data_ptr(), lda=lay.in_dim, a_layout=0, B=cps.w(l, e),
                                             ldb=lay.out_dim, b_layout=0, C=ubar[l][e].data_ptr(), ldc=lay.out_dim, M=B,
                                             N=lay.out_dim, K=lay.in_dim, epi=L.EPI_RELU_MASK, aux=run_c.H[l][e].data_ptr(),
                                             ldaux=lay.out_dim))
                    plan.add(f"G.ubar{l}", rt.gemm(fprobs, pick_cfg(B * E, lay.out_dim, rows_per_problem=B)))
                # dW_{l+1} += ubar_l^T v_{l+1};  for the head: dw_head += sum_b ubar_L  (v = ones)
                div_terms.append((l + 1, ubar[l], run_g.dZ[l + 1] if l + 1 < nh else None))
            return div_terms

        # The critics' forward on the data rows and the whole input-gradient chain of the diversity term read only the
        # batch and the (not yet updated) critics, so they run on a branch beside the policy-improvement step and the
        # target pass; the small-row launches emit no nested fork.
        par = B < TC_MIN_ROWS and os.environ.get("ORLK_EDAC_BRANCHES", "1") != "0"
        div_terms = None
        if par:
            plan.fork()
            plan.branch(1)
            data_forward()
            if self.eta > 0:
                div_terms = g_chain()
            plan.branch(0)
        actor_phase()
        tq = target_phase()
        if par:
            plan.join()
        else:
            data_forward()
        use_alpha = 0 if (pol._deterministic_backup or self.n_next > 1) else 1
        targs = (run_c.out.data_ptr(), B, E, tq.data_ptr(), B, E, lp_next.data_ptr(), self.scalars.data_ptr(), use_alpha,
                 self.rew.data_ptr(), self.term.data_ptr(), B, float(pol._gamma), run_c.dOut.data_ptr(), B, None,
                 self.loss_dev.data_ptr() + 4 * 12, self.loss_dev.data_ptr() + 4 * LS_TD_SUM)
        plan.add("C.td_loss", lambda: L.call("orlk_td_loss", *targs, rt.cur))
        emit_head_dgrad(rt, plan, run_c, "C.critics")
        emit_hidden_dgrad(rt, plan, run_c, "C.critics")
        # weight gradients of the TD loss: (layer, X [E,B,in] or a shared Mat, dY [E,B,out])
        td_terms = [(l, (mXd if l == 0 else run_c.H[l - 1]), (run_c.dZ[l] if l < nh else run_c.dOut)) for l in range(nh + 1)]
        self._emit_ens_wgrads(plan, "C.critics.wgrad_td", run_c, gb_c, td_terms, ones, split_base=0, with_bias=True,
                              k_splits=td_layout, tiny=(B < TC_MIN_ROWS and s_td == 1))
        if self.eta > 0:
            if div_terms is None:
                div_terms = g_chain()
            self._emit_ens_wgrads(plan, "G.wgrad_div", run_c, gb_c, div_terms, ones, split_base=s_td, with_bias=False,
                                  k_splits=None, tiny=B < TC_MIN_ROWS, col0_first=O)
        splits = [s_td + 1] * (nh + 1)
        plan.add("C.critics.adam", rt.adam(adam_descs(cps, gb_c, splits, polyak=True), self.groups_ptr))
        mask = (1 << self.g_actor) | (1 << self.g_c) | ((1 << self.g_alpha) if self.g_alpha >= 0 else 0)
        self.finish_ops(plan, mask)
        plan.keep.append(dict(locals()))       # runs, GradBufs, scratch: everything the closures point into
        self.plans["step"] = plan
        self._built = True

    def step(self, batch, noise=None) -> Dict[str, float]:
        self.bind_batch(batch)
        if not self._built:
            self._build()
        self.set_noise(noise)
        self.sync_lr()
        self.refresh()
        return self.result_of(self.run("step"))

    def result_of(self, out) -> Dict[str, float]:
        res = {"loss/actor": float(out[LS_ACTOR]),
               "loss/critics": float(out[LS_TD_SUM]) + (float(out[LS_DIV]) if self.eta > 0 else 0.0)}
        if self.auto_alpha:
            res["loss/alpha"] = float(out[LS_ALPHA_LOSS])
            res["alpha"] = float(out[LS_ALPHA])
        return res
