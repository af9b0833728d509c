"""Gradient-step engines for the twin-critic / tanh-Gaussian-actor algorithms: CQL and SAC (MOPO's learner).

Schedules follow the reference op order exactly:
  CQL  (policy/model_free/cql.py:87-207):  actor -> alpha -> [updated actor] TD target + conservative term
        (3-way logsumexp per repeat row) -> (Lagrange alpha') -> critic1, critic2 -> polyak.
  SAC  (policy/model_free/sac.py:88-140):  critics -> actor (updated critics) -> alpha (clamped) -> polyak.
What is *not* reproduced is the reference's wasted work: critic weight gradients in the actor step, actor
back-propagation from the critic losses, and the actor forward over the 10x repeated rows (the head is
evaluated once per distinct state and sampled 10 times).
"""
import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib as L
from .core import Mat, Plan
from .learner import (Learner, MlpRun, chainable, check_plain_mlp, emit_dact, emit_forward, emit_forward_pair, emit_head_dgrad, emit_hidden_dgrad,
                      emit_lo_refresh, emit_lo_refresh_many, emit_wgrad_adam, linears_of, make_gradbuf)
from .nets import TC_MIN_ROWS, ParamSet, dgrad_problem, pick_cfg

# loss block layout (floats)
LS_ACTOR, LS_ALPHA_LOSS, LS_ALPHA = 0, 1, 2
LS_C1, LS_C2, LS_CQL_ALPHA_LOSS, LS_CQL_ALPHA = 4, 5, 6, 7


class TwinCriticLearner(Learner):
    """Common state: actor ParamSet (fused mu|sigma head), twin critic ParamSet with targets, alpha scalars."""

    clamp_alpha = True

    def __init__(self, policy, batch_size: int):
        actor, c1, c2 = policy.actor, policy.critic1, policy.critic2
        super().__init__(actor.device)
        rt = self.rt
        self.policy, self.B = policy, int(batch_size)
        check_plain_mlp(actor.backbone, "actor")
        check_plain_mlp(c1.backbone, "critic")
        dist = actor.dist_net
        if (getattr(dist, "_sigma_min", -5.0), getattr(dist, "_sigma_max", 2.0)) != (-5.0, 2.0):
            # the sampler / backward kernels clamp log sigma to the reference's default [-5, 2] (dist_module.py:95-105)
            raise L.OrlkError("the CUDA engine implements TanhDiagGaussian with sigma_min=-5, sigma_max=2 only")
        if not (getattr(dist, "_c_sigma", False) and getattr(dist, "_unbounded", False)):
            raise L.OrlkError("SAC/CQL engine needs TanhDiagGaussian(unbounded=True, conditioned_sigma=True)")
        a_lin = linears_of(actor)
        self.actor_ps = ParamSet.from_linear_members(rt, "actor", [a_lin], fuse_last=2)
        self.critic_ps = ParamSet.from_linear_members(
            rt, "critics", [linears_of(c1), linears_of(c2)],
            targets=[linears_of(policy.critic1_old), linears_of(policy.critic2_old)], fuse_last=1)
        self.param_sets = [self.actor_ps, self.critic_ps]
        self.nh_a = len(self.actor_ps.layers) - 1
        self.nh_c = len(self.critic_ps.layers) - 1
        self.O = self.actor_ps.layers[0].in_dim
        self.A = self.actor_ps.layers[-1].out_dim // 2
        assert self.critic_ps.layers[0].in_dim == self.O + self.A

        tau = float(policy._tau)
        self.g_actor = self.add_group(policy.actor_optim)
        self.g_c1 = self.add_group(policy.critic1_optim, tau=tau)
        self.g_c2 = self.add_group(policy.critic2_optim, tau=tau)
        self.actor_ps.group_ids = [self.g_actor]
        self.critic_ps.group_ids = [self.g_c1, self.g_c2]
        self.auto_alpha = bool(policy._is_auto_alpha)
        self.alpha_mv = rt.zeros(2)
        self.g_alpha = -1
        if self.auto_alpha:
            self.g_alpha = self.add_group(policy.alpha_optim)
            self.target_entropy = float(policy._target_entropy)
            la = policy._log_alpha
            with torch.no_grad():
                self.scalars[L.SC_LOG_ALPHA] = float(la.detach().reshape(-1)[0])
                self.scalars[L.SC_ALPHA] = float(policy._alpha)
            la.data = self.scalars[L.SC_LOG_ALPHA:L.SC_LOG_ALPHA + 1].view(la.shape)
        else:
            self.target_entropy = 0.0
            with torch.no_grad():
                self.scalars[L.SC_ALPHA] = float(policy._alpha)
        self.gamma = float(policy._gamma)

    # ------------------------------------------------------------------ staging
    def _make_stage(self) -> None:
        rt, B, O, A = self.rt, self.B, self.O, self.A
        self.obs2 = rt.zeros(2 * B, O)
        self.act = rt.zeros(B, A)
        self.rew = rt.zeros(B, 1)
        self.term = rt.zeros(B, 1)
        self._bound_ptrs = None
        self._bound_token = None
        self._sources = None

    def bind_parts(self, parts) -> None:
        """Batches of several replay buffers (MBPolicyTrainer's real + model draws) as consecutive row blocks of the
        step's batch: the engine gathers each buffer's rows straight into its own staging, inside the step graph."""
        toks = tuple(p.token for p in parts)
        if self._sources is not None and tuple(t for t, _ in self._sources) == toks:
            return
        sizes = [int(t.idx_dev.shape[0]) for t in toks]
        if sum(sizes) != self.B:
            raise L.OrlkError(f"the step graph was built for batch size {self.B}, got parts {sizes}")
        offs = [sum(sizes[:i]) for i in range(len(sizes))]
        self._sources = list(zip(toks, offs))
        self._bound_token = None

    def gather_part_op(self, tok, row_off: int):
        B, O, A = self.B, self.O, self.A
        n = int(tok.idx_dev.shape[0])
        dst = (self.obs2.data_ptr() + 4 * row_off * O, self.obs2.data_ptr() + 4 * (B + row_off) * O,
               self.act.data_ptr() + 4 * row_off * A, self.rew.data_ptr() + 4 * row_off, self.term.data_ptr() + 4 * row_off)
        def op():
            tp, rows, row_w, o_dim, a_dim = tok.gather_args
            if o_dim != O or a_dim != A:
                raise L.OrlkError("replay buffer row layout does not match the policy's observation / action sizes")
            L.call("orlk_replay_gather_into", tp, rows, row_w, o_dim, a_dim, tok.idx_dev_ptr, n, *dst, self.rt.cur)
        return op

    def bind_batch(self, batch) -> None:
        """Use the replay buffer's persistent staging tensors as graph inputs (zero-copy), or copy a foreign
        batch into the engine's own staging."""
        if isinstance(batch, (list, tuple)):
            self.bind_parts(batch)
            return
        self._sources = None
        tok = getattr(batch, "token", None)
        if tok is not None and tok is self._bound_token:
            return              # the buffer's staging memory this engine's graphs are already bound to
        obs2 = getattr(batch, "obs2", None)
        if obs2 is not None and getattr(batch, "stable", False) and obs2.shape[0] == 2 * self.B:
            ptrs = (obs2.data_ptr(), batch["actions"].data_ptr(), batch["rewards"].data_ptr(),
                    batch["terminals"].data_ptr())
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
        with torch.no_grad():
            self.obs2[:B].copy_(torch.as_tensor(batch["observations"], device=self.dev, dtype=torch.float32))
            self.obs2[B:].copy_(torch.as_tensor(batch["next_observations"], device=self.dev, dtype=torch.float32))
            self.act.copy_(torch.as_tensor(batch["actions"], device=self.dev, dtype=torch.float32))
            self.rew.copy_(torch.as_tensor(batch["rewards"], device=self.dev, dtype=torch.float32).view(B, 1))
            self.term.copy_(torch.as_tensor(batch["terminals"], device=self.dev, dtype=torch.float32).view(B, 1))

    def set_noise(self, noise: Optional[Dict[str, torch.Tensor]]) -> None:
        """Parity mode: the caller supplies every random draw; otherwise Philox fills them."""
        if noise is None:
            self.set_noise_enabled(True)
            return
        self.set_noise_enabled(False)
        with torch.no_grad():
            for k, buf in self.noise_views.items():
                buf.copy_(torch.as_tensor(noise[k], device=self.dev, dtype=torch.float32).reshape(buf.shape))

    # ------------------------------------------------------------------ shared fragments
    def _emit_noise(self, plan: Plan, n_normal: int, n_uniform: int, lo: float, hi: float, defer: bool = False):
        """The step's noise block.  ``defer``: return the launch instead of adding it, so that the caller can put it on
        a branch parallel to work that does not consume noise."""
        args = (self.noise.data_ptr(), n_normal, n_uniform, lo, hi, int(self.seed), self.philox_counter.data_ptr(),
                self.noise_enable.data_ptr())
        op = ("philox", lambda: L.call("orlk_philox_fill", *args, self.rt.cur))
        if defer:
            return op
        plan.add(*op)

    def _emit_sample(self, plan: Plan, tag: str, head: torch.Tensor, head_row_off: int, rep: int, eps: torch.Tensor,
                     M: int, X: Mat, logp: torch.Tensor, obs: Mat) -> None:
        """a ~ pi(.|s) from head rows -> writes [obs | a] rows of the critic input X and logp."""
        O, A = self.O, self.A
        args = (head.data_ptr(), 2 * A, head_row_off, rep, eps.data_ptr(), M, A, X.ptr + 4 * O, X.ld, logp.data_ptr(),
                obs.ptr, obs.ld, O, X.ptr, X.ld)
        plan.add(tag, lambda: L.call("orlk_tanh_gauss_sample", *args, self.rt.cur))

    def _can_fuse_head_sample(self, run: MlpRun) -> bool:
        head = run.ps.layers[run.nh]
        if chainable(run, with_head=True):
            return False        # the head is the last stage of the pass's chain launch; the sampler follows on its own
        return (run.has_head and head.layout == "oi" and run.G == 1 and self.A <= 8 and head.in_dim % 4 == 0
                and run.M < TC_MIN_ROWS and os.environ.get("ORLK_FUSE_HEAD_SAMPLE", "1") != "0")

    def _emit_head_sample(self, plan: Plan, tag: str, run: MlpRun, uses) -> None:
        """Actor head + all the samplers that read it, one launch.  uses: (head_row0, head_row1, rep, eps, X, logp, obs)."""
        O, A = self.O, self.A
        head = run.ps.layers[run.nh]
        arr = (L.SampleUse * len(uses))()
        for i, (r0, r1, rep, eps, X, logp, obs) in enumerate(uses):
            u = arr[i]
            u.r0, u.r1, u.rep, u.obs_dim = r0, r1, rep, O
            u.eps, u.act, u.ld_act, u.logp = eps.data_ptr(), X.ptr + 4 * O, X.ld, logp.data_ptr()
            u.obs, u.ld_obs, u.xout, u.ld_x = obs.ptr, obs.ld, X.ptr, X.ld
        hin = run.H[run.nh - 1]
        args = (hin.data_ptr(), head.in_dim, run.ps.w(run.nh, 0, run.store), head.in_dim, run.ps.b(run.nh, 0, run.store),
                run.out.data_ptr(), run.M, head.in_dim, A, arr, len(uses))
        plan.keep.append(arr)
        plan.add(tag, lambda: L.call("orlk_head_sample", *args, self.rt.cur))

    def _emit_actor_forward(self, plan: Plan) -> None:
        """a ~ pi(.|s) for the policy-improvement step: actor forward + sampler, writes [s | a] into Xa and logp_a.
        Depends on the actor and the noise only, so SAC runs it beside the critics' backward pass."""
        rt, B = self.rt, self.B
        ar = self.run_actor
        obs = Mat.of(self.obs2).rows_(0, B)
        fuse_hs = self._can_fuse_head_sample(ar)
        emit_forward(rt, plan, ar, [obs], "A.actor", skip_head=fuse_hs)
        Xa = Mat.of(self.Xa)
        if fuse_hs:
            self._emit_head_sample(plan, "A.actor.head_sample", ar, [(0, B, 1, self.eps_actor, Xa, self.logp_a, obs)])
        else:
            self._emit_sample(plan, "A.sample", ar.out[0], 0, 1, self.eps_actor, B, Xa, self.logp_a, obs)

    def _emit_actor_update(self, plan: Plan, clamp01: bool, beside_forward=None, forward_done: bool = False) -> None:
        """a~pi(s); L = mean(alpha*logp - min Q); Adam(actor); alpha step.  (cql.py:93-106 / sac.py:111-126)
        ``beside_forward``: a (label, launch) that is independent of the actor forward (the noise fill) and runs on a
        parallel branch next to it.  ``forward_done``: the caller has already emitted ``_emit_actor_forward``."""
        rt, B, O, A = self.rt, self.B, self.O, self.A
        ar, cr = self.run_actor, self.run_critic_a
        obs = Mat.of(self.obs2).rows_(0, B)
        Xa = Mat.of(self.Xa)
        if not forward_done:
            if beside_forward is not None:
                plan.fork()
                plan.branch(1)
                if callable(beside_forward):
                    beside_forward()
                else:
                    plan.add(*beside_forward)
                plan.branch(0)
            fuse_hs = self._can_fuse_head_sample(ar)
            emit_forward(rt, plan, ar, [obs], "A.actor", skip_head=fuse_hs)
            if beside_forward is not None:
                plan.join()
            if fuse_hs:
                self._emit_head_sample(plan, "A.actor.head_sample", ar, [(0, B, 1, self.eps_actor, Xa, self.logp_a, obs)])
            else:
                self._emit_sample(plan, "A.sample", ar.out[0], 0, 1, self.eps_actor, B, Xa, self.logp_a, obs)
        q, dq = cr.out, cr.dOut      # [2, B, 1]
        chead = cr.ps.layers[cr.nh]
        forked = False
        fuse_th = (cr.G == 2 and cr.has_head and cr.NS == 1 and chead.layout == "oi" and not chainable(cr, with_head=True)
                   and not cr.fused_fwd and not cr.fuse_head_bwd and cr.dZT[cr.nh - 1] is None and B < TC_MIN_ROWS
                   and os.environ.get("ORLK_FUSE_TWIN_HEAD", "1") != "0")
        if fuse_th:
            # heads + d(loss)/dq + the heads' input gradient need no batch reduction: one launch on the critical path; the
            # loss value and the temperature step (which do) run beside the backward pass on q
            emit_forward(rt, plan, cr, [Xa, Xa], "A.critic", skip_head=True)
            K = chead.in_dim
            hargs = (cr.H[cr.nh - 1].data_ptr(), B * K, cr.ps.w(cr.nh, 0), chead.w_gs, cr.ps.b(cr.nh, 0), chead.b_gs,
                     self.scalars.data_ptr(), B, K, q.data_ptr(), B, dq.data_ptr(), B, self.glp.data_ptr(),
                     cr.dZ[cr.nh - 1].data_ptr(), B * K)
            plan.add("A.critic.head_fwd_bwd", lambda: L.call("orlk_twin_head_actor", *hargs, rt.cur))
            self._dq_scratch, self._glp_scratch = rt.zeros(2, B), rt.zeros(B)
            args = (q.data_ptr(), B, 2, self.logp_a.data_ptr(), B, self.scalars.data_ptr(), int(self.auto_alpha),
                    int(clamp01), self.target_entropy, self.groups_ptr, max(self.g_alpha, 0), self.alpha_mv.data_ptr(),
                    self._dq_scratch.data_ptr(), B, self._glp_scratch.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_ACTOR)
            plan.fork()
            plan.branch(1)
            plan.add("A.loss", lambda: L.call("orlk_sac_actor_loss", *args, rt.cur))
            plan.branch(0)
            forked = True
        else:
            emit_forward(rt, plan, cr, [Xa, Xa], "A.critic")
            args = (q.data_ptr(), B, 2, self.logp_a.data_ptr(), B, self.scalars.data_ptr(), int(self.auto_alpha),
                    int(clamp01), self.target_entropy, self.groups_ptr, max(self.g_alpha, 0), self.alpha_mv.data_ptr(),
                    dq.data_ptr(), B, self.glp.data_ptr(), self.loss_dev.data_ptr() + 4 * LS_ACTOR)
            plan.add("A.loss", lambda: L.call("orlk_sac_actor_loss", *args, rt.cur))
            emit_head_dgrad(rt, plan, cr, "A.critic")
        head = ar.out[0]
        c0, ah = cr.ps.layers[0], ar.ps.layers[ar.nh]
        fuse_entry = (not getattr(cr, "pending_head_dgrad", False) and c0.layout == "oi" and ah.layout == "oi" and ar.G == 1
                      and B < TC_MIN_ROWS and ar.dZT[ar.nh - 1] is None
                      and 4 * (cr.G * c0.out_dim * A + 2 * A * ah.in_dim) <= 48 * 1024
                      and os.environ.get("ORLK_FUSE_ACTOR_BWD", "1") != "0")
        if fuse_entry:
            # d/da through the critics' first layers, the sampler backward and the actor's head dgrad in ONE launch
            emit_hidden_dgrad(rt, plan, cr, "A.critic")
            eargs = (cr.dZ[0].data_ptr(), B * c0.out_dim, c0.out_dim, cr.G, cr.ps.w(0, 0), c0.w_gs, c0.in_dim, O,
                     head.data_ptr(), self.eps_actor.data_ptr(), Xa.ptr + 4 * O, Xa.ld, self.glp.data_ptr(), B, A,
                     ar.dOut.data_ptr(), ar.ps.w(ar.nh, 0), ah.in_dim, ar.H[ar.nh - 1].data_ptr(), ar.dZ[ar.nh - 1].data_ptr())
            plan.add("A.actor.bwd_entry", lambda: L.call("orlk_actor_bwd_entry", *eargs, rt.cur))
        else:
            # dL/da = sum over the two critics of dZ1 . W1[:, O:O+A]
            if getattr(cr, "pending_head_dgrad", False):       # fused chain: head dgrad, hidden dgrads and d/da in one launch
                emit_hidden_dgrad(rt, plan, cr, "A.critic", dact=(self.dA, O, A))
            else:
                emit_hidden_dgrad(rt, plan, cr, "A.critic")
                emit_dact(rt, plan, cr, self.dA, O, A, "A.critic")
            bargs = (head.data_ptr(), 2 * A, self.eps_actor.data_ptr(), Xa.ptr + 4 * O, Xa.ld, self.dA.data_ptr(), 2, B * A, A,
                     self.glp.data_ptr(), B, A, ar.dOut.data_ptr(), 2 * A)
            plan.add("A.head_bwd", lambda: L.call("orlk_tanh_gauss_bwd", *bargs, rt.cur))
            emit_head_dgrad(rt, plan, ar, "A.actor")
        nh = ar.nh
        split = (fuse_entry and nh >= 2 and not any(ar.tc_wgrad) and not chainable(ar, with_head=True)
                 and os.environ.get("ORLK_ACTOR_WGRAD_SPLIT", "1") != "0")
        if split:
            # weight gradients layer by layer, each as soon as its dZ exists (beside the remaining input-gradient launches);
            # a layer's Adam step only after its OWN input-gradient launch has read the old weights; the first layer - which
            # the next forward pass needs first - last, on the main stream
            wa = lambda tag, layers, **kw: emit_wgrad_adam(rt, plan, ar, [obs], self.gb_actor, self.groups_ptr, tag, polyak=False,
                                                           only_layers=layers, **kw)
            plan.fork()
            plan.branch(2)
            wa("A.actor.top", [nh - 1, nh], do_adam=False)
            plan.branch(0)
            pending = ("A.actor.top", [nh - 1, nh], 2)         # (tag, layers, side stream) whose Adam waits for the next dgrad
            for l in range(nh - 1, 0, -1):
                emit_hidden_dgrad(rt, plan, ar, "A.actor", down_to=l, from_layer=l)        # reads W[l], writes dZ[l - 1]
                plan.fork()
                plan.branch(pending[2])
                wa(pending[0], pending[1], do_wgrad=False)
                if l - 1 >= 1:
                    side = 3 if pending[2] == 2 else 2
                    plan.branch(side)
                    wa(f"A.actor.l{l - 1}", [l - 1], do_adam=False)
                    pending = (f"A.actor.l{l - 1}", [l - 1], side)
                plan.branch(0)
            wa("A.actor.l0", [0])
            plan.join()
            forked = False
        else:
            emit_hidden_dgrad(rt, plan, ar, "A.actor")
            emit_wgrad_adam(rt, plan, ar, [obs], self.gb_actor, self.groups_ptr, "A.actor", polyak=False)
        if forked:
            plan.join()

    def _alloc_actor_phase(self) -> None:
        rt, B, A = self.rt, self.B, self.A
        self.run_actor = self.mlp_run(self.actor_ps, B, self.nh_a, need_grad=True)
        self.run_critic_a = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=True)
        self.Xa = rt.zeros(B, self.O + A)
        self.logp_a = rt.zeros(B)
        self.glp = rt.zeros(B)
        self.dA = rt.zeros(2, B, A)

    def group_mask(self, *gs: int) -> int:
        m = 0
        for g in gs:
            if g >= 0:
                m |= 1 << g
        return m


class CQLLearner(TwinCriticLearner):
    """CQL, and COMBO's learner (policy/model_based/combo.py:109-243), which is the same step over the real+fake mix
    with two row ranges narrowed: ``n_real`` -- the leading rows whose Q enters the ``- w * mean Q`` term
    (combo.py:196-203) -- and ``cons_rows`` -- the rows the 3 x N conservative samples are drawn for (all of the mix,
    or the fake rows alone with rho_s="model", combo.py:162-166)."""

    def __init__(self, policy, batch_size: int, seed: int = 0, n_real: Optional[int] = None, cons_rows=None):
        super().__init__(policy, batch_size)
        rt = self.rt
        self.seed = seed
        self.N = int(policy._num_repeat_actions)
        # max_q_backup (cql.py:109-120): N next actions per row, each target critic maximised over them
        self.max_q_backup = bool(policy._max_q_backup)
        self.with_lagrange = bool(policy._with_lagrange)
        self.g_cql = -1
        self.cql_mv = rt.zeros(2)
        if self.with_lagrange:
            self.g_cql = self.add_group(policy.cql_alpha_optim)
            cla = policy.cql_log_alpha
            with torch.no_grad():
                self.scalars[L.SC_CQL_LOG_ALPHA] = float(cla.detach().reshape(-1)[0])
            cla.data = self.scalars[L.SC_CQL_LOG_ALPHA:L.SC_CQL_LOG_ALPHA + 1].view(cla.shape)
        self.act_lo = float(policy.action_space.low[0])
        self.act_hi = float(policy.action_space.high[0])
        self.push_groups()
        self._rebatch(self.B, n_real=n_real, cons_rows=cons_rows)

    def _rebatch(self, B: int, n_real: Optional[int] = None, cons_rows=None) -> None:
        rt, A = self.rt, self.A
        self.B = B
        self.n_real = B if n_real is None else int(n_real)
        self.c0, self.c1 = (0, B) if cons_rows is None else (int(cons_rows[0]), int(cons_rows[1]))
        if not (0 < self.n_real <= B and 0 <= self.c0 < self.c1 <= B):
            raise L.OrlkError("COMBO row split out of range")
        self.Bc = self.c1 - self.c0
        self.R = self.Bc * self.N
        self.Mc = B + 3 * self.R
        self.Rt = B * self.N if self.max_q_backup else B        # rows through the target critics
        self._make_stage()
        # noise block: normals first (eps_actor, eps_next, eps_pi, eps_pi_next), then uniforms (rand_act)
        R = self.R
        self.n_normal, self.n_uniform = (B + self.Rt + 2 * R) * A, R * A
        self.noise = rt.zeros(self.n_normal + self.n_uniform)
        o = 0
        views = {}
        for name, rows in (("eps_actor", B), ("eps_next", self.Rt), ("eps_pi", R), ("eps_pi_next", R), ("rand_act", R)):
            views[name] = self.noise[o:o + rows * A].view(rows, A)
            o += rows * A
        self.noise_views = views
        self.eps_actor = views["eps_actor"]
        self._built = False

    def _build(self) -> None:
        rt, B, O, A, R, Mc = self.rt, self.B, self.O, self.A, self.R, self.Mc
        self._alloc_actor_phase()
        self.run_actor_b = self.mlp_run(self.actor_ps, 2 * B, self.nh_a, need_grad=False)
        Rt, n_next = self.Rt, (self.N if self.max_q_backup else 1)
        # the target critics ride in the online critics' fused launch as a second job, whatever their row count (a pass of
        # their own on the small-row kernels would have to share the SMs with that launch)
        self.run_target = self.mlp_run(self.critic_ps, Rt, self.nh_c, need_grad=False, store="T", fused_min_rows=1)
        self.run_target.keep_h = False
        self.run_critic = self.mlp_run(self.critic_ps, Mc, self.nh_c, need_grad=True)
        if not self.run_critic.fused_fwd and Rt < TC_MIN_ROWS:
            self.run_target = self.mlp_run(self.critic_ps, Rt, self.nh_c, need_grad=False, store="T")
        self.Xt = rt.zeros(Rt, (O + A + 3) // 4 * 4)[:, :O + A]
        self.Xc = rt.zeros(Mc, (O + A + 3) // 4 * 4)[:, :O + A]      # 16-byte aligned rows: a TMA operand of the first layer
        self.lp_next, self.lp_pi, self.lp_pn = rt.zeros(Rt), rt.zeros(R), rt.zeros(R)
        self.loss_scratch = rt.zeros(L.load().orlk_cql_critic_loss_scratch_floats(B, R))
        self.gb_actor = make_gradbuf(rt, self.actor_ps, [self.run_actor])
        self.gb_critic = make_gradbuf(rt, self.critic_ps, [self.run_critic])

        plan = Plan(rt, "cql")
        noise = self._emit_noise(plan, self.n_normal, self.n_uniform, self.act_lo, self.act_hi, defer=True)

        def beside():
            # beside the actor forward: the noise fill and - the critics are only updated at the END of a step - the lo
            # words of the online / target critic weights that the fused critic passes fetch by TMA
            plan.add(*noise)
            items = ([(self.critic_ps, "P")] if self.run_critic.fused_fwd else []) \
                + ([(self.critic_ps, "T")] if self.run_target.fused_fwd else []) \
                + ([(self.critic_ps, "WT")] if self.run_critic.fused_bwd else [])
            emit_lo_refresh_many(rt, plan, items)
        self._emit_actor_update(plan, clamp01=False, beside_forward=beside)

        # ---- critic phase with the UPDATED actor (cql.py:108-192)
        obs2 = Mat.of(self.obs2)
        obs, nobs = obs2.rows_(0, B), obs2.rows_(B, 2 * B)
        c0, c1 = self.c0, self.c1
        cobs = obs.rows_(c0, c1)        # the states of the conservative term
        ab = self.run_actor_b
        fuse_hs = self._can_fuse_head_sample(ab)
        emit_forward(rt, plan, ab, [obs2], "C.actor", skip_head=fuse_hs)
        head = ab.out[0]
        Xt, Xc = Mat.of(self.Xt), Mat.of(self.Xc)
        v = self.noise_views
        # the samplers and the concat write disjoint row blocks of Xt / Xc: parallel branches
        plan.fork()
        plan.branch(1)
        if fuse_hs:     # one head pass feeds a'(s'), N x a(s) and N x a(s'): head + three samplers in one launch
            self._emit_head_sample(plan, "C.actor.head_sample", ab, [
                (B, 2 * B, n_next, v["eps_next"], Xt, self.lp_next, nobs),
                (c0, c1, self.N, v["eps_pi"], Xc.rows_(B, B + R), self.lp_pi, cobs),
                (B + c0, B + c1, self.N, v["eps_pi_next"], Xc.rows_(B + R, B + 2 * R), self.lp_pn, cobs)])
        else:
            self._emit_sample(plan, "C.sample_next", head, B, n_next, v["eps_next"], Rt, Xt, self.lp_next, nobs)
            plan.branch(2)
            self._emit_sample(plan, "C.sample_pi", head, c0, self.N, v["eps_pi"], R, Xc.rows_(B, B + R), self.lp_pi, cobs)
            plan.branch(3)
            self._emit_sample(plan, "C.sample_pi_next", head, B + c0, self.N, v["eps_pi_next"], R, Xc.rows_(B + R, B + 2 * R),
                              self.lp_pn, cobs)
        plan.branch(0)
        plan.add("C.concat", rt.concat([(Xc.rows_(0, B), obs, 1, Mat.of(self.act)),
                                        (Xc.rows_(B + 2 * R, Mc), cobs, self.N, Mat.of(v["rand_act"]))]))
        plan.join()
        # the target critics on (s', a') and the online critics on the 7936-row batch are independent: two branches
        cr = self.run_critic
        if not emit_forward_pair(rt, plan, cr, [Xc, Xc], "C.critic", self.run_target, [Xt, Xt], "C.target"):
            plan.fork()
            plan.branch(1)
            emit_forward(rt, plan, self.run_target, [Xt, Xt], "C.target")
            plan.branch(0)
            emit_forward(rt, plan, cr, [Xc, Xc], "C.critic")
            plan.join()
        pol = self.policy
        largs = (cr.out.data_ptr(), Mc, self.run_target.out.data_ptr(), Rt, self.lp_next.data_ptr(), self.lp_pi.data_ptr(),
                 self.lp_pn.data_ptr(), self.rew.data_ptr(), self.term.data_ptr(), B, self.n_real, n_next, R, A, self.gamma,
                 float(pol._cql_weight), float(pol._temperature), int(bool(pol._deterministic_backup)),
                 int(self.with_lagrange), float(pol._lagrange_threshold), self.scalars.data_ptr(), self.groups_ptr,
                 max(self.g_cql, 0), self.cql_mv.data_ptr(), cr.dOut.data_ptr(), Mc, self.loss_dev.data_ptr() + 4 * LS_C1,
                 self.loss_scratch.data_ptr())
        plan.add("C.loss", lambda: L.call("orlk_cql_critic_loss", *largs, rt.cur))
        self.emit_loss_readback(plan)       # every loss scalar is final: the copy overlaps the backward pass
        emit_head_dgrad(rt, plan, cr, "C.critic")
        emit_hidden_dgrad(rt, plan, cr, "C.critic")
        # (measured: the scalar head's SIMT weight gradient beside the input-gradient chain instead of beside the last
        # weight-gradient GEMM slows the chain more than it relieves the tail: 243 vs 235 us)
        emit_wgrad_adam(rt, plan, cr, [Xc, Xc], self.gb_critic, self.groups_ptr, "C.critic", polyak=True)
        self.finish_ops(plan, self.group_mask(self.g_actor, self.g_c1, self.g_c2, self.g_alpha, self.g_cql))
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
        res = {"loss/actor": float(out[LS_ACTOR]), "loss/critic1": float(out[LS_C1]), "loss/critic2": float(out[LS_C2])}
        if self.auto_alpha:
            res["loss/alpha"] = float(out[LS_ALPHA_LOSS])
            res["alpha"] = float(out[LS_ALPHA])
        if self.with_lagrange:
            res["loss/cql_alpha"] = float(out[LS_CQL_ALPHA_LOSS])
            res["cql_alpha"] = float(out[LS_CQL_ALPHA])
        return res


class SACLearner(TwinCriticLearner):
    """policy/model_free/sac.py:88-140 (MOPO's learner): critics -> actor (with the updated critics) -> alpha -> polyak.

    The polyak update is fused into the critics' Adam launch: the target networks are last read by the TD target
    earlier in the same step, so updating them right after the critic step is equivalent to the reference's
    end-of-step ``_sync_weight``."""

    def __init__(self, policy, batch_size: int, seed: int = 0):
        super().__init__(policy, batch_size)
        self.seed = seed
        self.push_groups()
        self._rebatch(self.B)

    def _rebatch(self, B: int) -> None:
        rt, A = self.rt, self.A
        self.B = B
        self._make_stage()
        self.n_normal = 2 * B * A
        self.noise = rt.zeros(self.n_normal)
        self.noise_views = {"eps_next": self.noise[:B * A].view(B, A), "eps_actor": self.noise[B * A:].view(B, A)}
        self.eps_actor = self.noise_views["eps_actor"]
        self._built = False

    def _build(self) -> None:
        rt, B, O, A = self.rt, self.B, self.O, self.A
        self._alloc_actor_phase()
        self.run_actor_n = self.mlp_run(self.actor_ps, B, self.nh_a, need_grad=False)
        self.run_target = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=False, store="T")
        self.run_critic = self.mlp_run(self.critic_ps, B, self.nh_c, need_grad=True)
        self.Xd, self.Xt = rt.zeros(B, O + A), rt.zeros(B, O + A)
        self.lp_next = rt.zeros(B)
        self.gb_actor = make_gradbuf(rt, self.actor_ps, [self.run_actor])
        self.gb_critic = make_gradbuf(rt, self.critic_ps, [self.run_critic])
        plan = Plan(rt, "sac")
        self._emit_noise(plan, self.n_normal, 0, 0.0, 1.0)
        obs2 = Mat.of(self.obs2)
        obs, nobs = obs2.rows_(0, B), obs2.rows_(B, 2 * B)
        Xd, Xt = Mat.of(self.Xd), Mat.of(self.Xt)
        cr = self.run_critic
        an = self.run_actor_n
        par = os.environ.get("ORLK_SAC_BRANCHES", "1") != "0"
        # Q(s, a_data) on one branch, a' ~ pi(s') and the target critics on another: they share no buffer
        if par:
            plan.fork()
            plan.branch(1)
        emit_forward(rt, plan, an, [nobs], "Q.actor_next")
        self._emit_sample(plan, "Q.sample_next", an.out[0], 0, 1, self.noise_views["eps_next"], B, Xt, self.lp_next, nobs)
        emit_forward(rt, plan, self.run_target, [Xt, Xt], "Q.target")
        if par:
            plan.branch(0)
        plan.add("Q.concat", rt.concat([(Xd, obs, 1, Mat.of(self.act))]))
        emit_forward(rt, plan, cr, [Xd, Xd], "Q.critic")
        if par:
            plan.join()
        targs = (cr.out.data_ptr(), B, 2, self.run_target.out.data_ptr(), B, 2, self.lp_next.data_ptr(),
                 self.scalars.data_ptr(), 1, self.rew.data_ptr(), self.term.data_ptr(), B, self.gamma, cr.dOut.data_ptr(), B,
                 None, self.loss_dev.data_ptr() + 4 * LS_C1, None)
        plan.add("Q.loss", lambda: L.call("orlk_td_loss", *targs, rt.cur))
        # the critics' backward pass and update on one branch, the policy-improvement step's actor forward + sampler
        # (which reads the actor and the noise only) on another; the critics it is scored by are the UPDATED ones
        if par:
            plan.fork()
            plan.branch(1)
            self._emit_actor_forward(plan)
            plan.branch(0)
        emit_head_dgrad(rt, plan, cr, "Q.critic")
        emit_hidden_dgrad(rt, plan, cr, "Q.critic")
        emit_wgrad_adam(rt, plan, cr, [Xd, Xd], self.gb_critic, self.groups_ptr, "Q.critic", polyak=True)
        if par:
            plan.join()
        self._emit_actor_update(plan, clamp01=True, forward_done=par)
        self.finish_ops(plan, self.group_mask(self.g_actor, self.g_c1, self.g_c2, self.g_alpha))
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
        res = {"loss/actor": float(out[LS_ACTOR]), "loss/critic1": float(out[LS_C1]), "loss/critic2": float(out[LS_C2])}
        if self.auto_alpha:
            res["loss/alpha"] = float(out[LS_ALPHA_LOSS])
            res["alpha"] = float(out[LS_ALPHA])
        return res
