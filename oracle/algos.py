"""CPU oracles for ``policy.learn`` of CQL / SAC(MOPO) / EDAC / IQL / TD3+BC.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Each class restates one
``learn`` of the reference on a flat parameter dict keyed like the reference's
``state_dict``.  Forward passes are written out (oracle/nets.py), backward is
torch autograd and the optimiser is ``torch.optim.Adam`` exactly as the
reference's run scripts construct it.  All random draws are passed in
(``noise``) in the order the reference consumes them, so
the CUDA engine, the oracle and the real reference can be fed identical noise.

After ``step`` every oracle exposes
  ``losses``  the reference's loss dict (python floats),
  ``grads``   name -> gradient tensor that entered the corresponding Adam step,
  ``state_dict()``  name -> parameter tensor (same keys as the reference).
"""
import math
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import nets

Tensors = Dict[str, torch.Tensor]


def _prefixed(state: Tensors, prefix: str) -> List[str]:
    return [k for k in state if k.startswith(prefix + ".")]


class _Learner:
    """Parameter store + Adam factories shared by all oracles."""

    device = "cpu"      # class-level default; set ``Cls.device = "cuda"`` before construction to time eager-GPU PyTorch

    def __init__(self, state: Tensors, trainable_prefixes: Iterable[str]):
        tp = tuple(trainable_prefixes)
        self.p: Tensors = {}
        for k, v in state.items():
            t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).detach().clone().to(self.device)
            if t.is_floating_point():
                t = t.float()
            train = any(k.startswith(pre + ".") for pre in tp) and "saved_" not in k and t.is_floating_point()
            self.p[k] = t.requires_grad_(train)
        self.grads: Tensors = {}
        self.losses: Dict[str, float] = {}

    def _adam(self, prefix: str, lr: float) -> torch.optim.Adam:
        names = _prefixed(self.p, prefix)
        opt = torch.optim.Adam([self.p[n] for n in names], lr=lr)
        opt.param_groups[0]["orlk_names"] = names      # (inside a param group: survives copy.deepcopy of the optimiser)
        return opt

    def _apply(self, opt: torch.optim.Adam, loss: torch.Tensor, retain_graph: bool = False) -> None:
        """zero_grad / backward / step, recording the gradients that Adam consumed."""
        opt.zero_grad()
        loss.backward(retain_graph=retain_graph)
        for n in opt.param_groups[0]["orlk_names"]:
            g = self.p[n].grad
            if g is not None:
                self.grads[n] = g.detach().clone()
        opt.step()

    def _polyak(self, old: str, new: str, tau: float) -> None:
        """sac.py:60-64 / edac.py:62-64 / iql.py:64-68 / td3bc.py:65-71."""
        for n in _prefixed(self.p, new):
            o = old + n[len(new):]
            if not self.p[o].is_floating_point():
                continue
            self.p[o].data.copy_(self.p[o].data * (1.0 - tau) + self.p[n].data * tau)

    def state_dict(self) -> Tensors:
        return {k: v.detach().clone() for k, v in self.p.items()}


class _AutoAlpha:
    """``alpha`` handling of sac.py:42-48,119-126 / cql.py:100-106 / edac.py:104-110."""

    def _init_alpha(self, alpha, clamp01: bool) -> None:
        self._clamp01 = clamp01
        if isinstance(alpha, tuple):
            self.auto_alpha = True
            self.target_entropy, log_alpha0, alpha_lr = alpha
            self.log_alpha = torch.tensor([float(log_alpha0)], requires_grad=True, device=self.device)
            self.alpha_optim = torch.optim.Adam([self.log_alpha], lr=alpha_lr)
            self.alpha = self.log_alpha.detach().exp()
        else:
            self.auto_alpha = False
            self.alpha = alpha

    def _alpha_update(self, log_probs: torch.Tensor) -> Optional[torch.Tensor]:
        if not self.auto_alpha:
            return None
        lp = log_probs.detach() + self.target_entropy
        alpha_loss = -(self.log_alpha * lp).mean()
        self.alpha_optim.zero_grad()
        alpha_loss.backward()
        self.grads["log_alpha"] = self.log_alpha.grad.detach().clone()
        self.alpha_optim.step()
        a = self.log_alpha.detach().exp()
        self.alpha = torch.clamp(a, 0.0, 1.0) if self._clamp01 else a
        return alpha_loss


class SACOracle(_Learner, _AutoAlpha):
    """policy/model_free/sac.py:88-140 (also MOPOPolicy.learn after the real/fake concat, mopo.py:81-84).

    ``alpha`` is a float or ``(target_entropy, log_alpha_init, alpha_lr)``.
    noise: ``eps_next`` [B,A] (sac.py:95), ``eps_actor`` [B,A] (sac.py:112).
    """

    def __init__(self, state, actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, alpha=0.2):
        super().__init__(state, ("actor", "critic1", "critic2"))
        self.tau, self.gamma = tau, gamma
        self.actor_optim = self._adam("actor", actor_lr)
        self.critic1_optim = self._adam("critic1", critic_lr)
        self.critic2_optim = self._adam("critic2", critic_lr)
        self._init_alpha(alpha, clamp01=True)

    def step(self, batch: Tensors, noise: Tensors) -> Dict[str, float]:
        p = self.p
        obs, act, nobs = batch["observations"], batch["actions"], batch["next_observations"]
        rew, term = batch["rewards"], batch["terminals"]
        self.grads = {}
        q1, q2 = nets.critic(p, "critic1", obs, act), nets.critic(p, "critic2", obs, act)
        with torch.no_grad():
            na, nlp = nets.actforward(p, "actor", nobs, noise["eps_next"])
            nq = torch.min(nets.critic(p, "critic1_old", nobs, na), nets.critic(p, "critic2_old", nobs, na))
            nq = nq - self.alpha * nlp
            target = rew + self.gamma * (1 - term) * nq
        c1 = (q1 - target).pow(2).mean()
        self._apply(self.critic1_optim, c1)
        c2 = (q2 - target).pow(2).mean()
        self._apply(self.critic2_optim, c2)

        a, lp = nets.actforward(p, "actor", obs, noise["eps_actor"])
        q1a, q2a = nets.critic(p, "critic1", obs, a), nets.critic(p, "critic2", obs, a)
        actor_loss = -torch.min(q1a, q2a).mean() + self.alpha * lp.mean()
        self._apply(self.actor_optim, actor_loss)
        alpha_loss = self._alpha_update(lp)
        self._polyak("critic1_old", "critic1", self.tau)
        self._polyak("critic2_old", "critic2", self.tau)

        out = {"loss/actor": actor_loss.item(), "loss/critic1": c1.item(), "loss/critic2": c2.item()}
        if self.auto_alpha:
            out["loss/alpha"] = alpha_loss.item()
            out["alpha"] = self.alpha.item()
        self.losses = out
        return out


class CQLOracle(_Learner, _AutoAlpha):
    """policy/model_free/cql.py:87-207.

    Keeps the reference's quirks: actor -> alpha -> critics
    order with the UPDATED actor in the critic phase; alpha not clamped; the
    discarded ``reshape`` at cql.py:153-157, so logsumexp runs over 3 columns
    for each of the B*N rows; the Lagrange multiplier's old value scales the
    critic losses.
    noise (consumption order): ``eps_actor`` [B,A] (:93), ``eps_next`` [B,A] (:124; [B*N,A]
    with max_q_backup, :113), ``rand_act`` [B*N,A] already scaled to the action box (:138-140),
    ``eps_pi`` [B*N,A] (:149), ``eps_pi_next`` [B*N,A] (:150).
    """

    def __init__(self, state, actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, alpha=0.2,
                 cql_weight=1.0, temperature=1.0, max_q_backup=False, deterministic_backup=True,
                 with_lagrange=True, lagrange_threshold=10.0, cql_alpha_lr=1e-4, num_repeat_actions=10):
        super().__init__(state, ("actor", "critic1", "critic2"))
        self.tau, self.gamma = tau, gamma
        self.actor_optim = self._adam("actor", actor_lr)
        self.critic1_optim = self._adam("critic1", critic_lr)
        self.critic2_optim = self._adam("critic2", critic_lr)
        self._init_alpha(alpha, clamp01=False)
        self.w, self.T = cql_weight, temperature
        self.max_q_backup, self.det_backup = max_q_backup, deterministic_backup
        self.with_lagrange, self.thr = with_lagrange, lagrange_threshold
        self.cql_log_alpha = torch.zeros(1, requires_grad=True, device=self.device)   # cql.py:57
        self.cql_alpha_optim = torch.optim.Adam([self.cql_log_alpha], lr=cql_alpha_lr)  # cql.py:58
        self.N = num_repeat_actions

    def step(self, batch: Tensors, noise: Tensors) -> Dict[str, float]:
        return self._step(batch, noise, batch, None)

    def _step(self, batch: Tensors, noise: Tensors, cons: Tensors, real) -> Dict[str, float]:
        """batch: the TD / actor rows; cons: the rows the conservative samples are drawn for; real: the rows of the
        conservative data term when they are not ``batch`` itself (COMBO)."""
        p, N = self.p, self.N
        obs, act, nobs = batch["observations"], batch["actions"], batch["next_observations"]
        rew, term = batch["rewards"], batch["terminals"]
        B = obs.shape[0]
        self.grads = {}

        # actor, cql.py:93-98
        a, lp = nets.actforward(p, "actor", obs, noise["eps_actor"])
        q1a, q2a = nets.critic(p, "critic1", obs, a), nets.critic(p, "critic2", obs, a)
        actor_loss = (self.alpha * lp - torch.min(q1a, q2a)).mean()
        self._apply(self.actor_optim, actor_loss)
        alpha_loss = self._alpha_update(lp)                                            # cql.py:100-106

        rep = lambda x: x.unsqueeze(1).repeat(1, N, 1).view(x.shape[0] * N, x.shape[-1])   # cql.py:142-147
        # TD target, cql.py:108-132
        with torch.no_grad():
            if self.max_q_backup:
                tn = rep(nobs)
                ta, _ = nets.actforward(p, "actor", tn, noise["eps_next"])
                t1 = nets.critic(p, "critic1_old", tn, ta).view(B, N, 1).max(1)[0].view(-1, 1)
                t2 = nets.critic(p, "critic2_old", tn, ta).view(B, N, 1).max(1)[0].view(-1, 1)
                nq = torch.min(t1, t2)
            else:
                na, nlp = nets.actforward(p, "actor", nobs, noise["eps_next"])
                nq = torch.min(nets.critic(p, "critic1_old", nobs, na), nets.critic(p, "critic2_old", nobs, na))
                if not self.det_backup:
                    nq = nq - self.alpha * nlp
        target = rew + self.gamma * (1 - term) * nq
        q1, q2 = nets.critic(p, "critic1", obs, act), nets.critic(p, "critic2", obs, act)
        c1 = (q1 - target).pow(2).mean()
        c2 = (q2 - target).pow(2).mean()

        # conservative term, cql.py:138-168
        rand_act = noise["rand_act"]
        tobs, tnobs = rep(cons["observations"]), rep(cons["next_observations"])
        a_pi, lp_pi = nets.actforward(p, "actor", tobs, noise["eps_pi"])             # calc_pi_values :62-72
        v1_pi = nets.critic(p, "critic1", tobs, a_pi) - lp_pi.detach()
        v2_pi = nets.critic(p, "critic2", tobs, a_pi) - lp_pi.detach()
        a_pn, lp_pn = nets.actforward(p, "actor", tnobs, noise["eps_pi_next"])
        v1_pn = nets.critic(p, "critic1", tobs, a_pn) - lp_pn.detach()
        v2_pn = nets.critic(p, "critic2", tobs, a_pn) - lp_pn.detach()
        log_u = math.log(0.5 ** rand_act.shape[-1])                                   # :82-83
        v1_r = nets.critic(p, "critic1", tobs, rand_act) - log_u
        v2_r = nets.critic(p, "critic2", tobs, rand_act) - log_u
        cat1 = torch.cat([v1_pi, v1_pn, v1_r], 1)                                     # [B*N, 3]  (:160-161)
        cat2 = torch.cat([v2_pi, v2_pn, v2_r], 1)
        if real is not None:                                                          # combo.py:196-197
            q1 = nets.critic(p, "critic1", real["observations"], real["actions"])
            q2 = nets.critic(p, "critic2", real["observations"], real["actions"])
        cons1 = torch.logsumexp(cat1 / self.T, dim=1).mean() * self.w * self.T - q1.mean() * self.w
        cons2 = torch.logsumexp(cat2 / self.T, dim=1).mean() * self.w * self.T - q2.mean() * self.w

        cql_alpha_loss = cql_alpha = None
        if self.with_lagrange:                                                        # :170-178
            cql_alpha = torch.clamp(self.cql_log_alpha.exp(), 0.0, 1e6)
            cons1 = cql_alpha * (cons1 - self.thr)
            cons2 = cql_alpha * (cons2 - self.thr)
            self.cql_alpha_optim.zero_grad()
            cql_alpha_loss = -(cons1 + cons2) * 0.5
            cql_alpha_loss.backward(retain_graph=True)
            self.grads["cql_log_alpha"] = self.cql_log_alpha.grad.detach().clone()
            self.cql_alpha_optim.step()

        c1 = c1 + cons1
        c2 = c2 + cons2
        self._apply(self.critic1_optim, c1, retain_graph=True)                        # :184-186
        self._apply(self.critic2_optim, c2)                                           # :188-190
        self._polyak("critic1_old", "critic1", self.tau)
        self._polyak("critic2_old", "critic2", self.tau)

        out = {"loss/actor": actor_loss.item(), "loss/critic1": c1.item(), "loss/critic2": c2.item()}
        if self.auto_alpha:
            out["loss/alpha"] = alpha_loss.item()
            out["alpha"] = self.alpha.item()
        if self.with_lagrange:
            out["loss/cql_alpha"] = cql_alpha_loss.item()
            out["cql_alpha"] = cql_alpha.item()
        self.losses = out
        return out


class COMBOOracle(CQLOracle):
    """policy/model_based/combo.py:109-243: the CQL step over the real+fake mix (combo.py:110-112).

    Differences from cql.py, all kept: the conservative samples are drawn for the mix rows, or for the fake rows only
    when ``rho_s == "model"`` (combo.py:162-166); the data term of the conservative loss is a fresh critic pass over
    the real rows (combo.py:196-203).  The discarded ``reshape`` (combo.py:183-187) is the same quirk as CQL's.
    noise: as CQLOracle, with ``eps_actor`` / ``eps_next`` over the mix rows and ``rand_act`` / ``eps_pi`` /
    ``eps_pi_next`` over (conservative rows) * N."""

    def __init__(self, state, rho_s="mix", **kw):
        super().__init__(state, **kw)
        self.rho_s = rho_s

    def step(self, batch, noise: Tensors) -> Dict[str, float]:
        real, fake = batch["real"], batch["fake"]
        mix = {k: torch.cat([real[k], fake[k]], 0) for k in real.keys()}
        return self._step(mix, noise, fake if self.rho_s == "model" else mix, real)


class EDACOracle(_Learner, _AutoAlpha):
    """policy/model_free/edac.py:88-166.  noise: ``eps_actor`` [B,A] (:96), ``eps_next`` [B,A] (:125), or
    [B*10,A] with max_q_backup (:117; the repeat count 10 is hard-coded there)."""

    def __init__(self, state, actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, alpha=0.2,
                 deterministic_backup=True, eta=1.0, max_q_backup=False):
        super().__init__(state, ("actor", "critics"))
        self.tau, self.gamma, self.eta = tau, gamma, eta
        self.det_backup, self.max_q_backup = deterministic_backup, max_q_backup
        self.actor_optim = self._adam("actor", actor_lr)
        self.critics_optim = self._adam("critics", critic_lr)
        self._init_alpha(alpha, clamp01=True)
        self.E = self.p["critics.model.0.weight"].shape[0]

    def step(self, batch: Tensors, noise: Tensors) -> Dict[str, float]:
        p, E = self.p, self.E
        obs, nobs = batch["observations"], batch["next_observations"]
        act = batch["actions"].detach().clone().requires_grad_(self.eta > 0)         # :92-93
        rew, term = batch["rewards"], batch["terminals"]
        self.grads = {}

        a, lp = nets.actforward(p, "actor", obs, noise["eps_actor"])
        qas = nets.ensemble_critic(p, "critics", obs, a)
        actor_loss = -torch.min(qas, 0)[0].mean() + self.alpha * lp.mean()
        self._apply(self.actor_optim, actor_loss)
        alpha_loss = self._alpha_update(lp)

        with torch.no_grad():
            if self.max_q_backup:                                                     # :113-122
                B = obs.shape[0]
                tn = nobs.unsqueeze(1).repeat(1, 10, 1).view(B * 10, nobs.shape[-1])
                ta, _ = nets.actforward(p, "actor", tn, noise["eps_next"])
                nq = nets.ensemble_critic(p, "critics_old", tn, ta).view(E, B, 10, 1).max(2)[0].view(E, B, 1).min(0)[0]
            else:
                na, nlp = nets.actforward(p, "actor", nobs, noise["eps_next"])
                nq = nets.ensemble_critic(p, "critics_old", nobs, na).min(0)[0]
                if not self.det_backup:
                    nq = nq - self.alpha * nlp
        target = rew + self.gamma * (1 - term) * nq
        qs = nets.ensemble_critic(p, "critics", obs, act)
        loss = (qs - target.unsqueeze(0)).pow(2).mean(dim=(1, 2)).sum()
        if self.eta > 0:                                                              # :136-149
            obs_t = obs.unsqueeze(0).repeat(E, 1, 1)
            act_t = act.unsqueeze(0).repeat(E, 1, 1).requires_grad_(True)
            q_t = nets.ensemble_critic(p, "critics", obs_t, act_t)
            g, = torch.autograd.grad(q_t.sum(), act_t, retain_graph=True, create_graph=True)
            g = g / (torch.norm(g, p=2, dim=2).unsqueeze(-1) + 1e-10)
            g = g.transpose(0, 1)
            gram = torch.einsum("bik,bjk->bij", g, g)
            mask = torch.eye(E, device=gram.device).unsqueeze(0).repeat(gram.size(0), 1, 1)
            gram = (1 - mask) * gram
            grad_loss = torch.mean(torch.sum(gram, dim=(1, 2))) / (E - 1)
            loss = loss + self.eta * grad_loss
        self._apply(self.critics_optim, loss)
        self._polyak("critics_old", "critics", self.tau)

        out = {"loss/actor": actor_loss.item(), "loss/critics": loss.item()}
        if self.auto_alpha:
            out["loss/alpha"] = alpha_loss.item()
            out["alpha"] = self.alpha.item()
        self.losses = out
        return out


class IQLOracle(_Learner):
    """policy/model_free/iql.py:86-139.  No random draws inside ``learn``."""

    def __init__(self, state, actor_lr=3e-4, critic_q_lr=3e-4, critic_v_lr=3e-4, tau=0.005, gamma=0.99,
                 expectile=0.8, temperature=0.1, max_mu=1.0):
        super().__init__(state, ("actor", "critic_q1", "critic_q2", "critic_v"))
        self.tau, self.gamma, self.expectile, self.temp, self.max_mu = tau, gamma, expectile, temperature, max_mu
        self.actor_optim = self._adam("actor", actor_lr)
        self.q1_optim = self._adam("critic_q1", critic_q_lr)
        self.q2_optim = self._adam("critic_q2", critic_q_lr)
        self.v_optim = self._adam("critic_v", critic_v_lr)

    def set_actor_lr(self, lr: float) -> None:
        """CosineAnnealingLR on the actor optimiser (run_iql.py:132-135) mutates param_groups."""
        for g in self.actor_optim.param_groups:
            g["lr"] = lr

    def step(self, batch: Tensors, noise: Tensors = None) -> Dict[str, float]:
        p = self.p
        obs, act, nobs = batch["observations"], batch["actions"], batch["next_observations"]
        rew, term = batch["rewards"], batch["terminals"]
        self.grads = {}
        with torch.no_grad():
            q = torch.min(nets.critic(p, "critic_q1_old", obs, act), nets.critic(p, "critic_q2_old", obs, act))
        v = nets.critic(p, "critic_v", obs)
        diff = q - v
        wgt = torch.where(diff > 0, self.expectile, 1 - self.expectile)               # :82-84
        v_loss = (wgt * diff ** 2).mean()
        self._apply(self.v_optim, v_loss)

        q1, q2 = nets.critic(p, "critic_q1", obs, act), nets.critic(p, "critic_q2", obs, act)
        with torch.no_grad():
            target = rew + self.gamma * (1 - term) * nets.critic(p, "critic_v", nobs)
        q1_loss = (q1 - target).pow(2).mean()
        q2_loss = (q2 - target).pow(2).mean()
        self._apply(self.q1_optim, q1_loss)
        self._apply(self.q2_optim, q2_loss)

        with torch.no_grad():
            q = torch.min(nets.critic(p, "critic_q1_old", obs, act), nets.critic(p, "critic_q2_old", obs, act))
            v = nets.critic(p, "critic_v", obs)
            exp_a = torch.clip(torch.exp((q - v) * self.temp), None, 100.0)
        mu, sigma = nets.gauss_head(p, "actor", obs, unbounded=False, max_mu=self.max_mu)
        logp = nets.normal_logp(act, mu, sigma)
        actor_loss = -(exp_a * logp).mean()
        self._apply(self.actor_optim, actor_loss)
        self._polyak("critic_q1_old", "critic_q1", self.tau)
        self._polyak("critic_q2_old", "critic_q2", self.tau)
        out = {"loss/actor": actor_loss.item(), "loss/q1": q1_loss.item(),
               "loss/q2": q2_loss.item(), "loss/v": v_loss.item()}
        self.losses = out
        return out


class TD3BCOracle(_Learner):
    """policy/model_free/td3bc.py:83-124.  noise: ``eps_target`` [B,A] = randn_like(actions) (:90)."""

    def __init__(self, state, actor_lr=3e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, max_action=1.0,
                 policy_noise=0.2, noise_clip=0.5, update_actor_freq=2, alpha=2.5):
        super().__init__(state, ("actor", "critic1", "critic2"))
        self.tau, self.gamma, self.max_action = tau, gamma, max_action
        self.policy_noise, self.noise_clip, self.freq, self.bc_alpha = policy_noise, noise_clip, update_actor_freq, alpha
        self.actor_optim = self._adam("actor", actor_lr)
        self.critic1_optim = self._adam("critic1", critic_lr)
        self.critic2_optim = self._adam("critic2", critic_lr)
        self.cnt, self.last_actor_loss = 0, 0

    def step(self, batch: Tensors, noise: Tensors) -> Dict[str, float]:
        p = self.p
        obs, act, nobs = batch["observations"], batch["actions"], batch["next_observations"]
        rew, term = batch["rewards"], batch["terminals"]
        self.grads = {}
        q1, q2 = nets.critic(p, "critic1", obs, act), nets.critic(p, "critic2", obs, act)
        with torch.no_grad():
            n = (noise["eps_target"] * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)
            na = (nets.det_actor(p, "actor_old", nobs, self.max_action) + n).clamp(-self.max_action, self.max_action)
            nq = torch.min(nets.critic(p, "critic1_old", nobs, na), nets.critic(p, "critic2_old", nobs, na))
            target = rew + self.gamma * (1 - term) * nq
        c1 = (q1 - target).pow(2).mean()
        c2 = (q2 - target).pow(2).mean()
        self._apply(self.critic1_optim, c1)
        self._apply(self.critic2_optim, c2)
        if self.cnt % self.freq == 0:                                                  # :107-116
            a = nets.det_actor(p, "actor", obs, self.max_action)
            q = nets.critic(p, "critic1", obs, a)
            lmbda = self.bc_alpha / q.abs().mean().detach()
            actor_loss = -lmbda * q.mean() + (a - act).pow(2).mean()
            self._apply(self.actor_optim, actor_loss)
            self.last_actor_loss = actor_loss.item()
            self._polyak("actor_old", "actor", self.tau)
            self._polyak("critic1_old", "critic1", self.tau)
            self._polyak("critic2_old", "critic2", self.tau)
        self.cnt += 1
        out = {"loss/actor": self.last_actor_loss, "loss/critic1": c1.item(), "loss/critic2": c2.item()}
        self.losses = out
        return out
