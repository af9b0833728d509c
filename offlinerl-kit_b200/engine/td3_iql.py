"""Gradient-step engines for TD3+BC (policy/model_free/td3bc.py:83-124) and IQL (policy/model_free/iql.py:86-139)."""
import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib as L
from .core import AdamT, Mat, Plan
from .learner import (Learner, MlpRun, check_plain_mlp, emit_dact, emit_forward, emit_head_dgrad, emit_hidden_dgrad, emit_wgrad_adam,
                      linears_of, make_gradbuf, polyak_descs)
from .nets import TC_MIN_ROWS, ParamSet, dgrad_problem

LS_ACTOR, LS_C1, LS_C2, LS_V = 0, 4, 5, 8


class _BatchMixin:
    """Batch staging shared by the engines in this file (same contract as TwinCriticLearner.bind_batch)."""

    def _make_stage(self) -> None:
        rt, B, O, A = self.rt, self.B, self.O, self.A
        self.obs2 = rt.zeros(2 * B, O)
        self.act = rt.zeros(B, A)
        self.rew = rt.zeros(B, 1)
        self.term = rt.zeros(B, 1)
        self._bound_ptrs = None
        self._bound_token = None

    def bind_batch(self, batch) -> None:
        tok = getattr(batch, "token", None)
        if tok is not None and tok is self._bound_token:
            return              # the buffer's staging memory this engine's graphs are already bound to (gather stays lazy)
        obs2 = getattr(batch, "obs2", None)
        if obs2 is not None and getattr(batch, "stable", False) and obs2.shape[0] == 2 * self.B:
            ptrs = (obs2.data_ptr(), batch["actions"].data_ptr(), batch["rewards"].data_ptr(), batch["terminals"].data_ptr())
            if self._bound_ptrs is None and not self.plans:
                self.obs2, self.act, self.rew, self.term = obs2, batch["actions"], batch["rewards"], batch["terminals"]
                self._bound_ptrs = ptrs
                self._bound_token = tok
                return
            if ptrs == self._bound_ptrs:
                self._bound_token = tok
                return
        # a foreign batch is copied into the staging memory; a still-pending draw of the bound buffer must not gather
        # over it at the head of the step graph
        self._bound_token = None
        B = self.B
        f32 = lambda x: torch.as_tensor(x, device=self.dev, dtype=torch.float32)
        with torch.no_grad():
            self.obs2[:B].copy_(f32(batch["observations"]))
            self.obs2[B:].copy_(f32(batch["next_observations"]))
            self.act.copy_(f32(batch["actions"]))
            self.rew.copy_(f32(batch["rewards"]).view(B, 1))
            self.term.copy_(f32(batch["terminals"]).view(B, 1))

    def set_noise(self, noise) -> None:
        if noise is None:
            self.set_noise_enabled(True)
            return
        self.set_noise_enabled(False)
        with torch.no_grad():
            for k, buf in self.noise_views.items():
                buf.copy_(torch.as_tensor(noise[k], device=self.dev, dtype=torch.float32).reshape(buf.shape))


class TD3BCLearner(_BatchMixin, Learner):
    """Twin critics every step; deterministic actor + BC term and the three polyak syncs every ``freq``-th step."""

    def __init__(self, policy, batch_size: int, seed: int = 0):
        actor = policy.actor
        super().__init__(actor.device)
        rt = self.rt
        self.policy, self.B, self.seed = policy, int(batch_size), seed
        check_plain_mlp(actor.backbone, "actor")
        check_plain_mlp(policy.critic1.backbone, "critic")
        self.actor_ps = ParamSet.from_linear_members(rt, "actor", [linears_of(actor)], targets=[linears_of(policy.actor_old)])
        self.critic_ps = ParamSet.from_linear_members(
            rt, "critics", [linears_of(policy.critic1), linears_of(policy.critic2)],
            targets=[linears_of(policy.critic1_old), linears_of(policy.critic2_old)])
        self.param_sets = [self.actor_ps, self.critic_ps]
        self.nh_a, self.nh_c = len(self.actor_ps.layers) - 1, len(self.critic_ps.layers) - 1
        self.O, self.A = self.actor_ps.layers[0].in_dim, self.actor_ps.layers[-1].out_dim
        tau = float(policy._tau)
        self.g_actor = self.add_group(policy.actor_optim, tau=tau)
        self.g_c1 = self.add_group(policy.critic1_optim, tau=tau)
        self.g_c2 = self.add_group(policy.critic2_optim, tau=tau)
        self.actor_ps.group_ids = [self.g_actor]
        self.critic_ps.group_ids = [self.g_c1, self.g_c2]
        self.push_groups()
        self._rebatch(self.B)

    def _rebatch(self, B: int) -> None:
        rt, A = self.rt, self.A
        self.B = B
        self._make_stage()
        self.noise = rt.zeros(B * A)
        self.noise_views = {"eps_target": self.noise.view(B, A)}
        self._built = False

    def _build(self) -> None:
        rt, B, O, A, pol = self.rt, self.B, self.O, self.A, self.policy
        max_a = float(pol._max_action)
        self.run_actor_t = self.mlp_run(self.actor_ps, B, self.nh_a, need_grad=False, store="T")
        self.run_target = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=False, store="T")
        self.run_critic = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=True)
        self.run_actor = self.mlp_run(self.actor_ps, B, self.nh_a, need_grad=True)
        self.run_q1 = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=True, members=1)
        self.Xd, self.Xt, self.Xa = rt.zeros(B, O + A), rt.zeros(B, O + A), rt.zeros(B, O + A)
        self.dA, self.dabc = rt.zeros(B, A), rt.zeros(B, A)
        self.gb_actor = make_gradbuf(rt, self.actor_ps, [self.run_actor])
        self.gb_critic = make_gradbuf(rt, self.critic_ps, [self.run_critic])
        obs2 = Mat.of(self.obs2)
        obs, nobs = obs2.rows_(0, B), obs2.rows_(B, 2 * B)
        Xd, Xt, Xa = Mat.of(self.Xd), Mat.of(self.Xt), Mat.of(self.Xa)

        # parallel graph branches for sub-chains that share no buffer (the small-row passes emit one launch per stage and
        # no nested fork; larger batches keep the sequential order)
        par = B < TC_MIN_ROWS and os.environ.get("ORLK_TD3_BRANCHES", "1") != "0"

        def critic_phase(plan: Plan, actor_forward=None) -> None:
            args = (self.noise.data_ptr(), B * A, 0, 0.0, 1.0, int(self.seed), self.philox_counter.data_ptr(),
                    self.noise_enable.data_ptr())
            plan.add("philox", lambda: L.call("orlk_philox_fill", *args, rt.cur))
            cr = self.run_critic
            at = self.run_actor_t
            # Q(s, a_data) on one branch, the target action and the target critics on another: no shared buffer
            if par:
                plan.fork()
                plan.branch(1)
            emit_forward(rt, plan, at, [nobs], "Q.actor_old")
            fargs = (at.out.data_ptr(), A, self.noise.data_ptr(), B, A, max_a, float(pol._policy_noise), float(pol._noise_clip),
                     Xt.ptr + 4 * O, Xt.ld, nobs.ptr, nobs.ld, O, Xt.ptr, Xt.ld)
            plan.add("Q.next_action", lambda: L.call("orlk_det_actor_fwd", *fargs, rt.cur))
            emit_forward(rt, plan, self.run_target, [Xt, Xt], "Q.target")
            if par:
                plan.branch(0)
            plan.add("Q.concat", rt.concat([(Xd, obs, 1, Mat.of(self.act))]))
            emit_forward(rt, plan, cr, [Xd, Xd], "Q.critic")
            if par:
                plan.join()
            targs = (cr.out.data_ptr(), B, 2, self.run_target.out.data_ptr(), B, 2, None, None, 0, self.rew.data_ptr(),
                     self.term.data_ptr(), B, float(pol._gamma), cr.dOut.data_ptr(), B, None,
                     self.loss_dev.data_ptr() + 4 * LS_C1, None)
            plan.add("Q.loss", lambda: L.call("orlk_td_loss", *targs, rt.cur))
            # the delayed policy update's actor forward reads the actor only: beside the critics' backward pass
            if actor_forward is not None and par:
                plan.fork()
                plan.branch(1)
                actor_forward(plan)
                plan.branch(0)
            emit_head_dgrad(rt, plan, cr, "Q.critic")
            emit_hidden_dgrad(rt, plan, cr, "Q.critic")
            emit_wgrad_adam(rt, plan, cr, [Xd, Xd], self.gb_critic, self.groups_ptr, "Q.critic", polyak=False)
            if actor_forward is not None and par:
                plan.join()

        def actor_forward(plan: Plan) -> None:
            ar = self.run_actor
            emit_forward(rt, plan, ar, [obs], "P.actor")
            fargs = (ar.out.data_ptr(), A, None, B, A, max_a, 0.0, 0.0, Xa.ptr + 4 * O, Xa.ld, obs.ptr, obs.ld, O, Xa.ptr, Xa.ld)
            plan.add("P.action", lambda: L.call("orlk_det_actor_fwd", *fargs, rt.cur))

        def actor_phase(plan: Plan) -> None:
            ar, q1 = self.run_actor, self.run_q1
            if not par:
                actor_forward(plan)
            emit_forward(rt, plan, q1, [Xa], "P.q1")
            largs = (q1.out.data_ptr(), Xa.ptr + 4 * O, Xa.ld, self.act.data_ptr(), A, B, A, float(pol._alpha),
                     q1.dOut.data_ptr(), self.dabc.data_ptr(), A, self.loss_dev.data_ptr() + 4 * LS_ACTOR)
            plan.add("P.loss", lambda: L.call("orlk_td3bc_actor_loss", *largs, rt.cur))
            emit_head_dgrad(rt, plan, q1, "P.q1")
            emit_hidden_dgrad(rt, plan, q1, "P.q1")
            emit_dact(rt, plan, q1, self.dA, O, A, "P.q1")
            bargs = (Xa.ptr + 4 * O, Xa.ld, self.dA.data_ptr(), A, self.dabc.data_ptr(), A, B, A, max_a, ar.dOut.data_ptr(), A)
            plan.add("P.head_bwd", lambda: L.call("orlk_det_actor_bwd", *bargs, rt.cur))
            emit_head_dgrad(rt, plan, ar, "P.actor")
            emit_hidden_dgrad(rt, plan, ar, "P.actor")
            emit_wgrad_adam(rt, plan, ar, [obs], self.gb_actor, self.groups_ptr, "P.actor", polyak=True)
            plan.add("P.polyak_critics", rt.adam(polyak_descs(self.critic_ps), self.groups_ptr))

        p_c = Plan(rt, "td3bc.critic")
        critic_phase(p_c)
        self.finish_ops(p_c, (1 << self.g_c1) | (1 << self.g_c2))
        p_ca = Plan(rt, "td3bc.critic+actor")
        critic_phase(p_ca, actor_forward)
        actor_phase(p_ca)
        self.finish_ops(p_ca, (1 << self.g_c1) | (1 << self.g_c2) | (1 << self.g_actor))
        self.plans = {"critic": p_c, "both": p_ca}
        self._built = True

    def step(self, batch, noise=None) -> Dict[str, float]:
        self.bind_batch(batch)
        if not self._built:
            self._build()
        self.set_noise(noise)
        self.sync_lr()
        self.refresh()
        key = self.next_key()
        self._many_keys, self._many_t = [key], 0
        return self.result_of(self.run(key))

    def next_key(self) -> str:
        """td3bc.py:107: the actor (and the three polyak syncs) every ``update_actor_freq``-th step"""
        pol = self.policy
        key = "both" if pol._cnt % pol._freq == 0 else "critic"
        pol._cnt += 1
        return key

    def result_of(self, out) -> Dict[str, float]:
        pol = self.policy
        if self._many_keys[self._many_t] == "both":
            pol._last_actor_loss = float(out[LS_ACTOR])
        return {"loss/actor": pol._last_actor_loss, "loss/critic1": float(out[LS_C1]), "loss/critic2": float(out[LS_C2])}


class IQLLearner(_BatchMixin, Learner):
    """V (expectile) -> Q1,Q2 (TD to r + gamma V(s')) -> actor (advantage-weighted log-likelihood) -> polyak(Q).

    q = min target-Q(s,a) is identical in the V phase and the actor phase (the targets only move at the end of the
    step), so it is computed once; the polyak update is fused into the Q critics' Adam launch for the same reason."""

    def __init__(self, policy, batch_size: int):
        actor = policy.actor
        super().__init__(actor.device)
        rt = self.rt
        self.policy, self.B = policy, int(batch_size)
        check_plain_mlp(actor.backbone, "actor")
        dist = actor.dist_net
        # (the tanh-squashed head is told by its class name: the reference's modules carry no flag for it, dist_module.py:79)
        squashed = any(c.__name__ == "TanhDiagGaussian" for c in type(dist).__mro__)
        if getattr(dist, "_c_sigma", True) or getattr(dist, "_unbounded", True) or squashed:
            raise L.OrlkError("IQL engine needs DiagGaussian(unbounded=False, conditioned_sigma=False)")
        self.max_mu = float(dist._max)
        self.actor_ps = ParamSet.from_linear_members(rt, "actor", [linears_of(actor)],
                                                     extra=[{"sigma_param": dist.sigma_param}])
        self.q_ps = ParamSet.from_linear_members(
            rt, "q", [linears_of(policy.critic_q1), linears_of(policy.critic_q2)],
            targets=[linears_of(policy.critic_q1_old), linears_of(policy.critic_q2_old)])
        self.v_ps = ParamSet.from_linear_members(rt, "v", [linears_of(policy.critic_v)])
        self.param_sets = [self.actor_ps, self.q_ps, self.v_ps]
        self.nh_a, self.nh_q, self.nh_v = (len(ps.layers) - 1 for ps in (self.actor_ps, self.q_ps, self.v_ps))
        self.O, self.A = self.actor_ps.layers[0].in_dim, self.actor_ps.layers[-1].out_dim
        tau = float(policy._tau)
        self.g_actor = self.add_group(policy.actor_optim)
        self.g_q1 = self.add_group(policy.critic_q1_optim, tau=tau)
        self.g_q2 = self.add_group(policy.critic_q2_optim, tau=tau)
        self.g_v = self.add_group(policy.critic_v_optim)
        self.actor_ps.group_ids, self.q_ps.group_ids, self.v_ps.group_ids = [self.g_actor], [self.g_q1, self.g_q2], [self.g_v]
        self.push_groups()
        self._rebatch(self.B)

    def _rebatch(self, B: int) -> None:
        self.B = B
        self._make_stage()
        self.noise_views = {}
        self._built = False

    def _build(self) -> None:
        rt, B, O, A, pol = self.rt, self.B, self.O, self.A, self.policy
        self.run_qt = self.mlp_run(self.q_ps, B, self.nh_q, need_grad=False, store="T")
        self.run_v = self.mlp_run(self.v_ps, B, self.nh_v, need_grad=True)
        self.run_q = self.mlp_run(self.q_ps, B, self.nh_q, need_grad=True)
        self.run_v2 = self.mlp_run(self.v_ps, 2 * B, self.nh_v, need_grad=False)
        self.run_actor = self.mlp_run(self.actor_ps, B, self.nh_a, need_grad=True)
        self.Xd = rt.zeros(B, O + A)
        self.qmin = rt.zeros(B)
        self.dsigma = rt.zeros(A)
        self.gb_actor = make_gradbuf(rt, self.actor_ps, [self.run_actor])
        self.gb_q = make_gradbuf(rt, self.q_ps, [self.run_q])
        self.gb_v = make_gradbuf(rt, self.v_ps, [self.run_v])
        obs2 = Mat.of(self.obs2)
        obs = obs2.rows_(0, B)
        Xd = Mat.of(self.Xd)
        plan = Plan(rt, "iql")
        plan.add("concat", rt.concat([(Xd, obs, 1, Mat.of(self.act))]))
        # Sub-chains that share no buffer run on parallel graph branches (small-row passes only: one launch per stage, no
        # nested fork).  The four forward passes that read only weights nobody has updated yet start together: the target
        # Q(s,a), V(s), the online Q(s,a) and the actor's mu(s).
        par = B < TC_MIN_ROWS and os.environ.get("ORLK_IQL_BRANCHES", "1") != "0"
        rv, rq, rv2, ra = self.run_v, self.run_q, self.run_v2, self.run_actor
        if par:
            plan.fork()
            plan.branch(1)
        # ---- V
        emit_forward(rt, plan, self.run_qt, [Xd, Xd], "V.qtarget")
        if par:
            plan.branch(2)
            emit_forward(rt, plan, rq, [Xd, Xd], "Q.q")
            plan.branch(3)
            emit_forward(rt, plan, ra, [obs], "P.actor")
            plan.branch(0)
        emit_forward(rt, plan, rv, [obs], "V.v")
        if par:
            plan.join()
        vargs = (self.run_qt.out.data_ptr(), B, rv.out.data_ptr(), B, float(pol._expectile), rv.dOut.data_ptr(),
                 self.qmin.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_V)
        plan.add("V.loss", lambda: L.call("orlk_iql_v_loss", *vargs, rt.cur))
        emit_head_dgrad(rt, plan, rv, "V.v")
        emit_hidden_dgrad(rt, plan, rv, "V.v")
        emit_wgrad_adam(rt, plan, rv, [obs], self.gb_v, self.groups_ptr, "V.v", polyak=False)
        # ---- Q (updated V on s' for the target and on s for the advantage, one 2B-row pass)
        if not par:
            emit_forward(rt, plan, rq, [Xd, Xd], "Q.q")
        emit_forward(rt, plan, rv2, [obs2], "Q.v_new")
        v_s, v_next = rv2.out.data_ptr(), rv2.out.data_ptr() + 4 * B
        targs = (rq.out.data_ptr(), B, 2, v_next, B, 1, None, None, 0, self.rew.data_ptr(), self.term.data_ptr(), B,
                 float(pol._gamma), rq.dOut.data_ptr(), B, None, self.loss_dev.data_ptr() + 4 * LS_C1, None)
        sp = self.actor_ps.extra_ptr("sigma_param")
        aargs = (ra.out.data_ptr(), A, sp, self.act.data_ptr(), A, self.qmin.data_ptr(), v_s, B, A, float(pol._temperature),
                 self.max_mu, ra.dOut.data_ptr(), A, self.dsigma.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_ACTOR)
        # the Q update and the actor update both start from V_new and touch disjoint parameters: two branches
        if par:
            plan.fork()
            plan.branch(1)
        else:
            emit_forward(rt, plan, ra, [obs], "P.actor")
        # ---- actor
        plan.add("P.loss", lambda: L.call("orlk_iql_actor_loss", *aargs, rt.cur))
        emit_head_dgrad(rt, plan, ra, "P.actor")
        emit_hidden_dgrad(rt, plan, ra, "P.actor")
        emit_wgrad_adam(rt, plan, ra, [obs], self.gb_actor, self.groups_ptr, "P.actor", polyak=False)
        if par:
            plan.branch(0)
        plan.add("Q.loss", lambda: L.call("orlk_td_loss", *targs, rt.cur))
        emit_head_dgrad(rt, plan, rq, "Q.q")
        emit_hidden_dgrad(rt, plan, rq, "Q.q")
        emit_wgrad_adam(rt, plan, rq, [Xd, Xd], self.gb_q, self.groups_ptr, "Q.q", polyak=True)
        if par:
            plan.join()
        off = self.actor_ps.extra["sigma_param"][0]
        ps = self.actor_ps
        plan.add("P.sigma_adam", rt.adam([AdamT(p=ps._ptr(ps.P, off), n=A, group=self.g_actor, m=ps._ptr(ps.Mo, off),
                                                v=ps._ptr(ps.Vo, off), grad=self.dsigma.data_ptr(), g_splits=1,
                                                g_split_stride=A)], self.groups_ptr))
        self.finish_ops(plan, (1 << self.g_actor) | (1 << self.g_q1) | (1 << self.g_q2) | (1 << self.g_v))
        self.plans["step"] = plan
        self._built = True

    def step(self, batch, noise=None) -> Dict[str, float]:
        self.bind_batch(batch)
        if not self._built:
            self._build()
        self.sync_lr()
        self.refresh()
        return self.result_of(self.run("step"))

    def result_of(self, out) -> Dict[str, float]:
        return {"loss/actor": float(out[LS_ACTOR]), "loss/q1": float(out[LS_C1]), "loss/q2": float(out[LS_C2]),
                "loss/v": float(out[LS_V])}
