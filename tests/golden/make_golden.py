"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the authoring container only (the reference tree is not on the GPU box):

    python tests/golden/make_golden.py

It imports the read-only reference from /root/reference (with the stub
gym/gymnasium/diffusers/wandb/matplotlib packages in tests/golden/_stubs), builds
each policy exactly as the corresponding run_example script does, overwrites
the parameters with the deterministic NumPy recipe of
``offlinerlkit_b200.synthetic.param_recipe`` (so the files stay small), runs
``policy.learn`` under a known torch seed, re-draws the same noise in the
reference's consumption order and stores
inputs / noise / losses / post-step parameters.  Before writing, it asserts
that the CPU oracle (oracle/) reproduces the reference on the same inputs, so
a golden file is only ever written from a run where oracle == reference.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

import gym
from offlinerlkit.nets import MLP
from offlinerlkit.modules import (ActorProb, Actor, Critic, EnsembleCritic, TanhDiagGaussian, DiagGaussian,
                                  EnsembleDynamicsModel)
from offlinerlkit.buffer import ReplayBuffer
from offlinerlkit.policy import CQLPolicy, EDACPolicy, IQLPolicy, TD3BCPolicy, SACPolicy, MOPOPolicy, COMBOPolicy
from offlinerlkit.dynamics import EnsembleDynamics
from offlinerlkit.utils.scaler import StandardScaler
from offlinerlkit.utils.termination_fns import (termination_fn_halfcheetah, termination_fn_hopper,
                                                termination_fn_walker2d)
from offlinerlkit.utils.noise import GaussianNoise

from offlinerlkit_b200.synthetic import make_dataset, param_recipe
from oracle import algos, replay as oreplay, dynamics as odyn

torch.set_num_threads(4)
LOSS_TOL = 2e-5      # oracle vs reference, relative (same ops on the same CPU: usually exact)


def overwrite_params(module: torch.nn.Module, seed: int) -> None:
    sd = module.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if v.is_floating_point()}
    vals = param_recipe(shapes, seed)
    with torch.no_grad():
        for k, v in vals.items():
            sd[k].copy_(torch.from_numpy(v))


def draw_batches(data, n_steps, batch, seed):
    """ReplayBuffer.sample of the reference under np.random.seed -> indices + gathered batches."""
    O, A = data["observations"].shape[1], data["actions"].shape[1]
    buf = ReplayBuffer(len(data["observations"]), (O,), np.float32, A, np.float32, device="cpu")
    buf.load_dataset(data)
    np.random.seed(seed)
    st = np.random.get_state()
    batches = [buf.sample(batch) for _ in range(n_steps)]
    np.random.set_state(st)
    idx = np.stack([oreplay.draw_indices(buf._size, batch) for _ in range(n_steps)])
    for t in range(n_steps):      # pin the oracle gather + index draw against the reference
        g = oreplay.gather({k: getattr(buf, k) for k in oreplay.FIELDS}, idx[t])
        for k in oreplay.FIELDS:
            assert np.array_equal(g[k], batches[t][k].numpy()), k
    return idx, batches


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def check_losses(ref, ora, tag):
    assert ref.keys() == ora.keys(), (tag, ref.keys(), ora.keys())
    for k in ref:
        assert abs(ref[k] - ora[k]) <= LOSS_TOL * max(1.0, abs(ref[k])), (tag, k, ref[k], ora[k])


def check_state(ref_sd, ora_sd, tag, tol=2e-5):
    for k, v in ref_sd.items():
        if not v.is_floating_point():
            continue
        r = rel(ora_sd[k].numpy(), v.numpy())
        assert r <= tol, (tag, k, r)


def tensor_stats(sd):
    """Compact, order-sensitive fingerprint of a state dict for the full-size configs."""
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point():
            continue
        x = v.detach().double().flatten().numpy()
        stride = max(1, x.size // 64)
        out[k] = np.concatenate([[x.sum(), np.abs(x).sum(), np.sqrt((x * x).sum())], x[::stride][:64]])
    return out


class GradTap:
    """Records what every Adam step of the REFERENCE consumed: a step pre-hook on each of the reference's own optimisers
    copies ``p.grad`` of its parameters (name as in ``state_dict``; bare tensors like ``log_alpha`` by the given name).
    ``p.grad`` read after ``learn`` would be wrong for CQL's actor (the critic losses back-propagate into it later)."""

    def __init__(self, module, optims, extra=None):
        self.names = {id(p): n for n, p in module.named_parameters()}
        for n, t in (extra or {}).items():
            self.names[id(t)] = n
        self.grads = {}
        for opt in optims:
            opt.register_step_pre_hook(self._hook)

    def _hook(self, opt, args, kwargs):
        for grp in opt.param_groups:
            for p in grp["params"]:
                if p.grad is not None:
                    self.grads[self.names[id(p)]] = p.grad.detach().clone()

    def take(self):
        g, self.grads = self.grads, {}
        return g


def grad_stats(grads):
    """Per-tensor gradient fingerprint: [sum, abs-sum, L2, max-abs, 64 strided elements] (float64)."""
    out = {}
    for k, v in grads.items():
        x = v.detach().double().flatten().numpy()
        stride = max(1, x.size // 64)
        out[k] = np.concatenate([[x.sum(), np.abs(x).sum(), np.sqrt((x * x).sum()), np.abs(x).max()], x[::stride][:64]])
    return out


def check_grads(ref_g, ora_g, tag, tol=2e-5):
    """oracle.grads == the gradients the reference's optimisers consumed (relative to each tensor's largest element)."""
    assert set(ref_g) == set(ora_g), (tag, sorted(set(ref_g) ^ set(ora_g)))
    for k, v in ref_g.items():
        r = rel(ora_g[k].numpy(), v.numpy())
        assert r <= tol, (tag, "grad", k, r)


def pack_grads(store, t, ref_g, full_state):
    pack(store, f"gradstats{t}", grad_stats(ref_g))
    if full_state:
        pack(store, f"grads{t}", ref_g)


def pack(store, prefix, d):
    for k, v in d.items():
        store[f"{prefix}|{k}"] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)


def save(name, store, meta, full_state):
    store["meta"] = np.array(json.dumps(meta))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KB  (full_state={full_state})")


# ----------------------------------------------------------------------------------------------
def build_sac_like(O, A, hidden, device="cpu"):
    actor_backbone = MLP(input_dim=O, hidden_dims=hidden)
    c1b = MLP(input_dim=O + A, hidden_dims=hidden)
    c2b = MLP(input_dim=O + A, hidden_dims=hidden)
    dist = TanhDiagGaussian(latent_dim=actor_backbone.output_dim, output_dim=A, unbounded=True, conditioned_sigma=True)
    return ActorProb(actor_backbone, dist, device), Critic(c1b, device), Critic(c2b, device)


def gen_cql(name, O, A, hidden, B, N, n_steps, with_lagrange, full_state, det_backup=True, n_data=4096, seed=0,
            max_q_backup=False):
    data = make_dataset(n_data, O, A, seed=0)
    torch.manual_seed(seed)
    actor, c1, c2 = build_sac_like(O, A, hidden)
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 100 + i)
    hyper = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, cql_weight=5.0, temperature=1.0,
                 max_q_backup=max_q_backup, deterministic_backup=det_backup, with_lagrange=with_lagrange,
                 lagrange_threshold=10.0, cql_alpha_lr=3e-4, num_repeat_actions=N)
    alpha_lr, target_entropy = 1e-4, -A
    log_alpha = torch.zeros(1, requires_grad=True)
    pol = CQLPolicy(actor, c1, c2,
                    torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                    torch.optim.Adam(c1.parameters(), lr=hyper["critic_lr"]),
                    torch.optim.Adam(c2.parameters(), lr=hyper["critic_lr"]),
                    action_space=gym.spaces.Box(-1, 1, (A,)), tau=hyper["tau"], gamma=hyper["gamma"],
                    alpha=(target_entropy, log_alpha, torch.optim.Adam([log_alpha], lr=alpha_lr)),
                    cql_weight=hyper["cql_weight"], temperature=hyper["temperature"],
                    max_q_backup=max_q_backup, deterministic_backup=det_backup, with_lagrange=with_lagrange,
                    lagrange_threshold=hyper["lagrange_threshold"], cql_alpha_lr=hyper["cql_alpha_lr"],
                    num_repeart_actions=N)
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critic1_optim, pol.critic2_optim, pol.alpha_optim, pol.cql_alpha_optim],
                  {"log_alpha": log_alpha, "cql_log_alpha": pol.cql_log_alpha})
    pre = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    ora = algos.CQLOracle(pre, alpha=(target_entropy, 0.0, alpha_lr), **hyper)
    idx, batches = draw_batches(data, n_steps, B, seed=seed)
    store = {"idx": idx}
    R = B * N
    for t in range(n_steps):
        torch.manual_seed(1000 + t)
        ref_loss = pol.learn(batches[t])
        ref_g = tap.take()
        torch.manual_seed(1000 + t)
        noise = {"eps_actor": torch.randn(B, A), "eps_next": torch.randn(R if max_q_backup else B, A),
                 "rand_act": torch.FloatTensor(R, A).uniform_(-1.0, 1.0),
                 "eps_pi": torch.randn(R, A), "eps_pi_next": torch.randn(R, A)}
        ora_loss = ora.step(batches[t], noise)
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"noise{t}", noise)
        pack(store, f"loss{t}", {k: np.float64(v) for k, v in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    store["log_alpha_final"] = log_alpha.detach().numpy()
    store["cql_log_alpha_final"] = pol.cql_log_alpha.detach().numpy()
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="cql", O=O, A=A, hidden=hidden, B=B, N=N, n_steps=n_steps, n_data=n_data, data_seed=0,
                param_seeds={"actor": 100, "critic1": 101, "critic2": 102}, alpha_lr=alpha_lr,
                target_entropy=target_entropy, hyper=hyper, np_seed=seed)
    save(name, store, meta, full_state)


def gen_combo(name, O, A, hidden, n_real, n_fake, N, n_steps, rho_s, with_lagrange, full_state, det_backup=True,
              n_data=4096, seed=0):
    """COMBOPolicy.learn (combo.py:109-243) on {"real": ..., "fake": ...} batches; the fake rows are draws from a second
    synthetic dataset (what the model buffer holds is irrelevant to the step's arithmetic)."""
    data, fdata = make_dataset(n_data, O, A, seed=0), make_dataset(n_data, O, A, seed=7)
    torch.manual_seed(seed)
    actor, c1, c2 = build_sac_like(O, A, hidden)
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 100 + i)
    hyper = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, cql_weight=5.0, temperature=1.0,
                 max_q_backup=False, deterministic_backup=det_backup, with_lagrange=with_lagrange,
                 lagrange_threshold=10.0, cql_alpha_lr=3e-4, num_repeat_actions=N)
    alpha_lr, target_entropy = 1e-4, -A
    log_alpha = torch.zeros(1, requires_grad=True)
    pol = COMBOPolicy(None, actor, c1, c2,
                      torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                      torch.optim.Adam(c1.parameters(), lr=hyper["critic_lr"]),
                      torch.optim.Adam(c2.parameters(), lr=hyper["critic_lr"]),
                      action_space=gym.spaces.Box(-1, 1, (A,)), tau=hyper["tau"], gamma=hyper["gamma"],
                      alpha=(target_entropy, log_alpha, torch.optim.Adam([log_alpha], lr=alpha_lr)),
                      cql_weight=hyper["cql_weight"], temperature=hyper["temperature"],
                      max_q_backup=False, deterministic_backup=det_backup, with_lagrange=with_lagrange,
                      lagrange_threshold=hyper["lagrange_threshold"], cql_alpha_lr=hyper["cql_alpha_lr"],
                      num_repeart_actions=N, uniform_rollout=False, rho_s=rho_s)
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critic1_optim, pol.critic2_optim, pol.alpha_optim, pol.cql_alpha_optim],
                  {"log_alpha": log_alpha, "cql_log_alpha": pol.cql_log_alpha})
    pre = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    ora = algos.COMBOOracle(pre, rho_s=rho_s, alpha=(target_entropy, 0.0, alpha_lr), **hyper)
    idx, batches = draw_batches(data, n_steps, n_real, seed=seed)
    fidx, fbatches = draw_batches(fdata, n_steps, n_fake, seed=seed + 1)
    store = {"idx": idx, "fake_idx": fidx}
    B = n_real + n_fake
    R = (n_fake if rho_s == "model" else B) * N
    for t in range(n_steps):
        both = {"real": batches[t], "fake": fbatches[t]}
        torch.manual_seed(1000 + t)
        ref_loss = pol.learn(both)
        ref_g = tap.take()
        torch.manual_seed(1000 + t)
        noise = {"eps_actor": torch.randn(B, A), "eps_next": torch.randn(B, A),
                 "rand_act": torch.FloatTensor(R, A).uniform_(-1.0, 1.0),
                 "eps_pi": torch.randn(R, A), "eps_pi_next": torch.randn(R, A)}
        ora_loss = ora.step(both, noise)
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"noise{t}", noise)
        pack(store, f"loss{t}", {k: np.float64(v) for k, v in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    store["log_alpha_final"] = log_alpha.detach().numpy()
    store["cql_log_alpha_final"] = pol.cql_log_alpha.detach().numpy()
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="combo", O=O, A=A, hidden=hidden, B=B, n_real=n_real, n_fake=n_fake, N=N, rho_s=rho_s, n_steps=n_steps,
                n_data=n_data, data_seed=0, fake_data_seed=7, param_seeds={"actor": 100, "critic1": 101, "critic2": 102},
                alpha_lr=alpha_lr, target_entropy=target_entropy, hyper=hyper, np_seed=seed)
    save(name, store, meta, full_state)


def gen_sac(name, O, A, hidden, B, n_steps, full_state, n_data=4096, seed=0):
    data = make_dataset(n_data, O, A, seed=0)
    actor, c1, c2 = build_sac_like(O, A, hidden)
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 110 + i)
    hyper = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99)
    alpha_lr, target_entropy = 1e-4, -A
    log_alpha = torch.zeros(1, requires_grad=True)
    pol = SACPolicy(actor, c1, c2,
                    torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                    torch.optim.Adam(c1.parameters(), lr=hyper["critic_lr"]),
                    torch.optim.Adam(c2.parameters(), lr=hyper["critic_lr"]),
                    tau=hyper["tau"], gamma=hyper["gamma"],
                    alpha=(target_entropy, log_alpha, torch.optim.Adam([log_alpha], lr=alpha_lr)))
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critic1_optim, pol.critic2_optim, pol.alpha_optim], {"log_alpha": log_alpha})
    pre = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    ora = algos.SACOracle(pre, alpha=(target_entropy, 0.0, alpha_lr), **hyper)
    idx, batches = draw_batches(data, n_steps, B, seed=seed)
    store = {"idx": idx}
    for t in range(n_steps):
        torch.manual_seed(2000 + t)
        ref_loss = pol.learn(batches[t])
        ref_g = tap.take()
        torch.manual_seed(2000 + t)
        noise = {"eps_next": torch.randn(B, A), "eps_actor": torch.randn(B, A)}
        ora_loss = ora.step(batches[t], noise)
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"noise{t}", noise)
        pack(store, f"loss{t}", {k: np.float64(v) for k, v in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    store["log_alpha_final"] = log_alpha.detach().numpy()
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="sac", O=O, A=A, hidden=hidden, B=B, n_steps=n_steps, n_data=n_data, data_seed=0,
                param_seeds={"actor": 110, "critic1": 111, "critic2": 112}, alpha_lr=alpha_lr,
                target_entropy=target_entropy, hyper=hyper, np_seed=seed)
    save(name, store, meta, full_state)


def gen_edac(name, O, A, hidden, E, B, n_steps, full_state, eta=1.0, n_data=4096, seed=0, max_q_backup=False):
    data = make_dataset(n_data, O, A, seed=0)
    actor_backbone = MLP(input_dim=O, hidden_dims=hidden)
    dist = TanhDiagGaussian(latent_dim=actor_backbone.output_dim, output_dim=A, unbounded=True, conditioned_sigma=True)
    actor = ActorProb(actor_backbone, dist, "cpu")
    critics = EnsembleCritic(O, A, hidden, num_ensemble=E, device="cpu")
    overwrite_params(actor, 120)
    overwrite_params(critics, 121)
    hyper = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, deterministic_backup=False, eta=eta, max_q_backup=max_q_backup)
    alpha_lr, target_entropy = 1e-4, -A
    log_alpha = torch.zeros(1, requires_grad=True)
    pol = EDACPolicy(actor, critics, torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                     torch.optim.Adam(critics.parameters(), lr=hyper["critic_lr"]),
                     tau=hyper["tau"], gamma=hyper["gamma"],
                     alpha=(target_entropy, log_alpha, torch.optim.Adam([log_alpha], lr=alpha_lr)),
                     max_q_backup=max_q_backup, deterministic_backup=False, eta=eta)
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critics_optim, pol.alpha_optim], {"log_alpha": log_alpha})
    pre = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    ora = algos.EDACOracle(pre, alpha=(target_entropy, 0.0, alpha_lr), **hyper)
    idx, batches = draw_batches(data, n_steps, B, seed=seed)
    store = {"idx": idx}
    for t in range(n_steps):
        torch.manual_seed(3000 + t)
        ref_loss = pol.learn({k: v.clone() for k, v in batches[t].items()})
        ref_g = tap.take()
        torch.manual_seed(3000 + t)
        noise = {"eps_actor": torch.randn(B, A), "eps_next": torch.randn(B * 10 if max_q_backup else B, A)}
        ora_loss = ora.step(batches[t], noise)
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"noise{t}", noise)
        pack(store, f"loss{t}", {k: np.float64(v) for k, v in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    store["log_alpha_final"] = log_alpha.detach().numpy()
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="edac", O=O, A=A, hidden=hidden, E=E, B=B, n_steps=n_steps, n_data=n_data, data_seed=0,
                param_seeds={"actor": 120, "critics": 121}, alpha_lr=alpha_lr, target_entropy=target_entropy,
                hyper=hyper, np_seed=seed)
    save(name, store, meta, full_state)


def gen_iql(name, O, A, hidden, B, n_steps, full_state, n_data=4096, seed=0):
    data = make_dataset(n_data, O, A, seed=0)
    ab = MLP(input_dim=O, hidden_dims=hidden, dropout_rate=None)
    dist = DiagGaussian(latent_dim=ab.output_dim, output_dim=A, unbounded=False, conditioned_sigma=False)
    actor = ActorProb(ab, dist, "cpu")
    q1, q2 = Critic(MLP(O + A, hidden), "cpu"), Critic(MLP(O + A, hidden), "cpu")
    v = Critic(MLP(O, hidden), "cpu")
    for i, m in enumerate((actor, q1, q2, v)):
        overwrite_params(m, 130 + i)
    hyper = dict(actor_lr=3e-4, critic_q_lr=3e-4, critic_v_lr=3e-4, tau=0.005, gamma=0.99, expectile=0.7,
                 temperature=3.0)
    pol = IQLPolicy(actor, q1, q2, v,
                    torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                    torch.optim.Adam(q1.parameters(), lr=hyper["critic_q_lr"]),
                    torch.optim.Adam(q2.parameters(), lr=hyper["critic_q_lr"]),
                    torch.optim.Adam(v.parameters(), lr=hyper["critic_v_lr"]),
                    action_space=gym.spaces.Box(-1, 1, (A,)), tau=hyper["tau"], gamma=hyper["gamma"],
                    expectile=hyper["expectile"], temperature=hyper["temperature"])
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critic_q1_optim, pol.critic_q2_optim, pol.critic_v_optim])
    pre = {k: v_.detach().clone() for k, v_ in pol.state_dict().items()}
    ora = algos.IQLOracle(pre, **hyper)
    idx, batches = draw_batches(data, n_steps, B, seed=seed)
    store = {"idx": idx}
    for t in range(n_steps):
        ref_loss = pol.learn(batches[t])
        ref_g = tap.take()
        ora_loss = ora.step(batches[t])
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"loss{t}", {k: np.float64(v_) for k, v_ in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="iql", O=O, A=A, hidden=hidden, B=B, n_steps=n_steps, n_data=n_data, data_seed=0,
                param_seeds={"actor": 130, "critic_q1": 131, "critic_q2": 132, "critic_v": 133}, hyper=hyper,
                np_seed=seed)
    save(name, store, meta, full_state)


def gen_td3bc(name, O, A, hidden, B, n_steps, full_state, n_data=4096, seed=0):
    data = make_dataset(n_data, O, A, seed=0)
    actor = Actor(MLP(O, hidden), A, device="cpu")
    c1, c2 = Critic(MLP(O + A, hidden), "cpu"), Critic(MLP(O + A, hidden), "cpu")
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 140 + i)
    hyper = dict(actor_lr=3e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, max_action=1.0, policy_noise=0.2,
                 noise_clip=0.5, update_actor_freq=2, alpha=2.5)
    pol = TD3BCPolicy(actor, c1, c2,
                      torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                      torch.optim.Adam(c1.parameters(), lr=hyper["critic_lr"]),
                      torch.optim.Adam(c2.parameters(), lr=hyper["critic_lr"]),
                      tau=hyper["tau"], gamma=hyper["gamma"], max_action=1.0,
                      exploration_noise=GaussianNoise(sigma=0.1), policy_noise=0.2, noise_clip=0.5,
                      update_actor_freq=2, alpha=2.5, scaler=None)
    pol.train()
    tap = GradTap(pol, [pol.actor_optim, pol.critic1_optim, pol.critic2_optim])
    pre = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    ora = algos.TD3BCOracle(pre, **hyper)
    idx, batches = draw_batches(data, n_steps, B, seed=seed)
    store = {"idx": idx}
    for t in range(n_steps):
        torch.manual_seed(4000 + t)
        ref_loss = pol.learn(batches[t])
        ref_g = tap.take()
        torch.manual_seed(4000 + t)
        noise = {"eps_target": torch.randn(B, A)}
        ora_loss = ora.step(batches[t], noise)
        check_losses(ref_loss, ora_loss, f"{name} step {t}")
        check_state(pol.state_dict(), ora.state_dict(), f"{name} step {t}")
        pack(store, f"noise{t}", noise)
        pack(store, f"loss{t}", {k: np.float64(v) for k, v in ref_loss.items()})
        check_grads(ref_g, ora.grads, f"{name} step {t}")
        pack_grads(store, t, ref_g, full_state)
        pack(store, f"stats{t}", tensor_stats(pol.state_dict()))
    if full_state:
        pack(store, "post", pol.state_dict())
    meta = dict(algo="td3bc", O=O, A=A, hidden=hidden, B=B, n_steps=n_steps, n_data=n_data, data_seed=0,
                param_seeds={"actor": 140, "critic1": 141, "critic2": 142}, hyper=hyper, np_seed=seed)
    save(name, store, meta, full_state)


def gen_dynamics(name, O, A, hidden, E, n_elites, B, n_batches, S, full_state, term="halfcheetah", seed=0):
    """EnsembleDynamics.learn (n_batches mini-batches), validate, and one imagination step of S states."""
    wds = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4][:len(hidden)] + [1e-4]
    model = EnsembleDynamicsModel(O, A, hidden, num_ensemble=E, num_elites=n_elites, weight_decays=wds, device="cpu")
    overwrite_params(model, 150)
    with torch.no_grad():   # keep the learned log-variance bounds at their reference init (dynamics_module.py:69-76)
        model.max_logvar.fill_(0.5)
        model.min_logvar.fill_(-10.0)
    optim = torch.optim.Adam(model.parameters(), lr=1e-3)
    tfn = {"halfcheetah": termination_fn_halfcheetah, "hopper": termination_fn_hopper,
           "walker2d": termination_fn_walker2d}[term]
    ofn = {"halfcheetah": odyn.term_halfcheetah, "hopper": odyn.term_hopper, "walker2d": odyn.term_walker2d}[term]
    data = make_dataset(4096, O, A, seed=0)
    data["rewards"] = data["rewards"].reshape(-1, 1)          # as ReplayBuffer.sample_all returns them
    dyn = EnsembleDynamics(model, optim, StandardScaler(), tfn, penalty_coef=0.5)
    inputs, targets = dyn.format_samples_for_training(data)
    dyn.scaler.fit(inputs)
    mu, std = odyn.scaler_fit(inputs)
    assert np.array_equal(mu, dyn.scaler.mu) and np.array_equal(std, dyn.scaler.std)
    x = dyn.scaler.transform(inputs)
    rng = np.random.default_rng(7)
    boot = rng.integers(0, len(x), size=(E, B * n_batches))
    xin, yin = x[boot], targets[boot]
    pre = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ora = odyn.DynamicsOracle(pre, wds, lr=1e-3)
    tap = GradTap(model, [optim])
    ref_loss = dyn.learn(xin, yin, batch_size=B, logvar_loss_coef=0.01)
    ref_g = tap.take()                      # what Adam consumed in the LAST mini-batch
    ora_loss = ora.learn(xin, yin, batch_size=B, logvar_loss_coef=0.01)
    assert abs(ref_loss - ora_loss) <= LOSS_TOL * max(1, abs(ref_loss)), (ref_loss, ora_loss)
    check_grads(ref_g, ora.grads, name)
    check_state(model.state_dict(), {k: v.detach() for k, v in ora.p.items()}, name)
    hold = slice(0, 256)
    ref_val = dyn.validate(x[hold], targets[hold])
    ora_val = ora.validate(x[hold], targets[hold])
    assert rel(ora_val, ref_val) < 1e-5
    # imagination step: replay the two NumPy draws of ensemble_dynamics.py:48 / dynamics_module.py:118
    obs, act = data["observations"][:S].copy(), data["actions"][:S].copy()
    if term != "halfcheetah":
        obs[:, 0] = 1.0 + 0.05 * obs[:, 0]
        obs[:, 1] = 0.05 * obs[:, 1]
    np.random.seed(11)
    st = np.random.get_state()
    r_nobs, r_rew, r_term, r_info = dyn.step(obs, act)
    np.random.set_state(st)
    noise = np.random.normal(size=(E, S, O + 1))
    midx = np.random.choice(model.elites.data.cpu().numpy(), size=S)
    o_nobs, o_rew, o_term, o_info = ora.step(obs, act, mu, std, ofn, 0.5, noise, midx)
    assert rel(o_nobs, r_nobs) < 1e-5 and rel(o_rew, r_rew) < 1e-5 and np.array_equal(o_term, r_term)
    assert rel(o_info["penalty"], r_info["penalty"]) < 1e-5
    other = {}
    for mode in ("pairwise-diff", "ensemble_std"):       # the other two penalties of ensemble_dynamics.py:60-70
        dyn._uncertainty_mode = mode
        np.random.set_state(st)
        _, m_rew, _, m_info = dyn.step(obs, act)
        _, o_rew, _, o_info2 = ora.step(obs, act, mu, std, ofn, 0.5, noise, midx, uncertainty_mode=mode)
        assert rel(o_info2["penalty"], m_info["penalty"]) < 1e-5 and rel(o_rew, m_rew) < 1e-5, mode
        other["step_penalty_" + mode] = m_info["penalty"]
        other["step_reward_" + mode] = m_rew
    dyn._uncertainty_mode = "aleatoric"
    store = {**other, "boot": boot, "learn_loss": np.float64(ref_loss), "val": np.asarray(ref_val, np.float64),
             "step_obs": obs, "step_act": act, "step_noise": noise.astype(np.float64), "step_midx": midx,
             "step_next_obs": r_nobs, "step_reward": r_rew, "step_terminal": r_term,
             "step_penalty": r_info["penalty"], "step_raw_reward": r_info["raw_reward"],
             "scaler_mu": mu, "scaler_std": std}
    pack(store, "stats", tensor_stats(model.state_dict()))
    pack(store, "gradstats_last", grad_stats(ref_g))
    if full_state:
        pack(store, "post", model.state_dict())
        pack(store, "grads_last", ref_g)
    meta = dict(algo="dynamics", O=O, A=A, hidden=hidden, E=E, n_elites=n_elites, B=B, n_batches=n_batches, S=S,
                weight_decays=wds, lr=1e-3, term=term, penalty_coef=0.5, param_seed=150, boot_seed=7, data_seed=0,
                n_data=4096, holdout=256)
    save(name, store, meta, full_state)


def gen_sample_next(name, O, A, hidden, E, n_elites, S, num_samples, elites):
    """EnsembleDynamics.sample_next_obss (ensemble_dynamics.py:81-99), the first half of MOBILE's penalty: the torch CPU
    generator is seeded, so the reference's ``torch.randn_like(std)`` draws can be replayed as explicit noise."""
    wds = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4][:len(hidden)] + [1e-4]
    model = EnsembleDynamicsModel(O, A, hidden, num_ensemble=E, num_elites=n_elites, weight_decays=wds, device="cpu")
    overwrite_params(model, 170)
    with torch.no_grad():
        model.max_logvar.fill_(0.5)
        model.min_logvar.fill_(-10.0)
    model.set_elites(list(elites))
    data = make_dataset(2048, O, A, seed=0)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
    inputs, _ = dyn.format_samples_for_training(data)
    dyn.scaler.fit(inputs)
    mu, std = odyn.scaler_fit(inputs)
    obs, act = data["observations"][:S].copy(), data["actions"][:S].copy()
    torch.manual_seed(21)
    ref = dyn.sample_next_obss(torch.as_tensor(obs), torch.as_tensor(act), num_samples).numpy()
    torch.manual_seed(21)
    noise = torch.stack([torch.randn(len(elites), S, O + 1) for _ in range(num_samples)], 0).numpy()
    pre = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ora = odyn.DynamicsOracle(pre, wds, lr=1e-3)
    got = odyn.sample_next_obss(ora, obs, act, mu, std, np.asarray(elites), noise)
    assert ref.shape == (num_samples, len(elites), S, O) and rel(got, ref) < 1e-5, (ref.shape, rel(got, ref))
    store = {"obs": obs, "act": act, "noise": noise, "next_obss": ref, "scaler_mu": mu, "scaler_std": std,
             "elites": np.asarray(elites, np.int64)}
    meta = dict(algo="dynamics_sample_next", O=O, A=A, hidden=hidden, E=E, n_elites=n_elites, S=S, num_samples=num_samples,
                weight_decays=wds, param_seed=170, data_seed=0, n_data=2048, torch_seed=21)
    save(name, store, meta, True)


def gen_rollout(name, O, A, hidden, dyn_hidden, E, n_elites, S, horizon, term="hopper", uniform=False):
    """MOPOPolicy.rollout (mopo.py:45-79): compaction order and per-step RNG consumption.
    uniform: COMBOPolicy.rollout with uniform_rollout=True (combo.py:67-107) -- NumPy uniform actions, no actor."""
    wds = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4][:len(dyn_hidden)] + [1e-4]
    model = EnsembleDynamicsModel(O, A, dyn_hidden, num_ensemble=E, num_elites=n_elites, weight_decays=wds, device="cpu")
    overwrite_params(model, 160)
    with torch.no_grad():
        model.max_logvar.fill_(0.5)
        model.min_logvar.fill_(-10.0)
        model.output_layer.weight.mul_(0.1)             # small deltas / small std so that states
        model.output_layer.bias[..., O + 1:] = -6.0     # terminate gradually over the horizon
    tfn = {"halfcheetah": termination_fn_halfcheetah, "hopper": termination_fn_hopper,
           "walker2d": termination_fn_walker2d}[term]
    data = make_dataset(2048, O, A, seed=0)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), tfn, penalty_coef=0.5)
    inputs, _ = dyn.format_samples_for_training(data)
    dyn.scaler.fit(inputs)
    actor, c1, c2 = build_sac_like(O, A, hidden)
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 161 + i)
    opts = (torch.optim.Adam(actor.parameters(), lr=1e-4), torch.optim.Adam(c1.parameters(), lr=3e-4),
            torch.optim.Adam(c2.parameters(), lr=3e-4))
    if uniform:
        pol = COMBOPolicy(dyn, actor, c1, c2, *opts, action_space=gym.spaces.Box(-1, 1, (A,)), alpha=0.2,
                          uniform_rollout=True)
    else:
        pol = MOPOPolicy(dyn, actor, c1, c2, *opts, alpha=0.2)
    init = data["observations"][:S].copy()
    init[:, 0] = 1.0 + 0.3 * init[:, 0]       # heights around the hopper/walker2d thresholds -> some terminate
    init[:, 1] = 0.1 * init[:, 1]
    torch.manual_seed(77)
    np.random.seed(78)
    st_t, st_n = torch.get_rng_state(), np.random.get_state()
    out, info = pol.rollout(init, horizon)
    # replay the draws: per step eps [S_t, A] (torch), normal [E, S_t, D] then choice (numpy)
    torch.set_rng_state(st_t)
    np.random.set_state(st_n)
    counts, eps_l, nrm_l, mid_l = [], [], [], []
    n_done, S_t = 0, S
    while n_done < len(out["obss"]):
        if uniform:
            eps_l.append(np.random.uniform(-1.0, 1.0, size=(S_t, A)))       # combo.py:82-86
        else:
            eps_l.append(torch.randn(S_t, A).numpy())
        nrm_l.append(np.random.normal(size=(E, S_t, O + 1)))
        mid_l.append(np.random.choice(model.elites.data.cpu().numpy(), size=S_t))
        counts.append(S_t)
        term_t = out["terminals"][n_done:n_done + S_t]
        n_done += S_t
        S_t = int((~term_t).sum())
    assert n_done == len(out["obss"]) == info["num_transitions"], (n_done, len(out["obss"]))
    if uniform:
        assert np.array_equal(np.concatenate(eps_l), out["actions"]), "uniform action stream differs from the reference"
    store = {"init": init, "counts": np.asarray(counts), "scaler_mu": dyn.scaler.mu, "scaler_std": dyn.scaler.std,
             ("uniform_actions" if uniform else "eps"): np.concatenate(eps_l), "normal": np.concatenate([n.reshape(E, -1) for n in nrm_l], axis=1),
             "midx": np.concatenate(mid_l), "reward_mean": np.float64(info["reward_mean"])}
    for k, v in out.items():
        store["out|" + k] = v
    pack(store, "dyn", model.state_dict())
    pack(store, "actor", actor.state_dict())
    meta = dict(algo="rollout", O=O, A=A, hidden=hidden, dyn_hidden=dyn_hidden, E=E, n_elites=n_elites, S=S,
                horizon=horizon, term=term, penalty_coef=0.5, weight_decays=wds, uniform=uniform)
    save(name, store, meta, True)


class _StubLogger:
    """What EnsembleDynamics.train needs of utils/logger.py:Logger: log / logkv / set_timestep / dumpkvs / model_dir."""

    def __init__(self, model_dir):
        self.model_dir, self.rows, self._kv = model_dir, [], {}

    def log(self, *a, **k):
        pass

    def logkv(self, k, v):
        self._kv[k] = float(v)

    def set_timestep(self, t):
        self._kv["timestep"] = t

    def dumpkvs(self, exclude=None):
        self.rows.append(dict(self._kv))
        self._kv = {}


def gen_dynamics_train(name, O, A, hidden, E, n_elites, n_data, B, max_epochs, max_epochs_since_update=5, seed=5):
    """EnsembleDynamics.train (ensemble_dynamics.py:111-176) end to end: holdout split, scaler, bootstrap matrix, epochs with
    update_save / early stop, elites, load_save."""
    import tempfile
    wds = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4][:len(hidden)] + [1e-4]
    model = EnsembleDynamicsModel(O, A, hidden, num_ensemble=E, num_elites=n_elites, weight_decays=wds, device="cpu")
    overwrite_params(model, 170)
    with torch.no_grad():
        model.max_logvar.fill_(0.5)
        model.min_logvar.fill_(-10.0)
    pre = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data = make_dataset(n_data, O, A, seed=3)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    # a learnable target: next_obs = obs + a smooth function of (obs, act), so that the holdout loss really improves
    rng = np.random.default_rng(11)
    Wd = rng.standard_normal((O + A, O)).astype(np.float32) * 0.3
    xin = np.concatenate([data["observations"], data["actions"]], 1)
    data["next_observations"] = (data["observations"] + np.tanh(xin @ Wd) + 0.05 * data["next_observations"]).astype(np.float32)
    data["rewards"] = (np.sin(xin.sum(1, keepdims=True)) + 0.05 * data["rewards"]).astype(np.float32)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_hopper,
                           penalty_coef=0.5)
    kw = dict(max_epochs=max_epochs, max_epochs_since_update=max_epochs_since_update, batch_size=B, holdout_ratio=0.2,
              logvar_loss_coef=0.01)
    with tempfile.TemporaryDirectory() as tmp:
        logger = _StubLogger(tmp)
        torch.manual_seed(seed)
        np.random.seed(seed + 1)
        dyn.train(data, logger, **kw)
    ora = odyn.DynamicsOracle(pre, wds, lr=1e-3)
    inputs, targets = dyn.format_samples_for_training(data)
    torch.manual_seed(seed)
    np.random.seed(seed + 1)
    res = odyn.train(ora, inputs, targets, n_elites, **kw)
    assert res["epochs"] == len(logger.rows), (res["epochs"], len(logger.rows))
    for row, (tl, hl, _) in zip(logger.rows, res["log"]):
        assert abs(row["loss/dynamics_train_loss"] - tl) <= LOSS_TOL * max(1, abs(tl)), (row, tl)
        assert abs(row["loss/dynamics_holdout_loss"] - hl) <= LOSS_TOL * max(1, abs(hl)), (row, hl)
    assert res["elites"] == model.elites.data.tolist(), (res["elites"], model.elites.data.tolist())
    assert np.array_equal(res["mu"], dyn.scaler.mu) and np.array_equal(res["std"], dyn.scaler.std)
    sd = model.state_dict()
    check_state({k: v for k, v in sd.items() if k != "elites"}, {k: v.detach() for k, v in ora.p.items()}, name)
    store = {"train_loss": np.asarray([r["loss/dynamics_train_loss"] for r in logger.rows], np.float64),
             "holdout_loss": np.asarray([r["loss/dynamics_holdout_loss"] for r in logger.rows], np.float64),
             "member_holdout": np.asarray([m for _, _, m in res["log"]], np.float64),
             "elites": np.asarray(res["elites"]), "scaler_mu": dyn.scaler.mu, "scaler_std": dyn.scaler.std,
             "Wd": Wd}
    pack(store, "post", {k: v for k, v in sd.items() if k != "elites"})
    meta = dict(algo="dynamics", O=O, A=A, hidden=hidden, E=E, n_elites=n_elites, n_data=n_data, data_seed=3, B=B,
                weight_decays=wds, lr=1e-3, param_seed=170, torch_seed=seed, np_seed=seed + 1, train_kw=kw,
                epochs=len(logger.rows), term="hopper")
    save(name, store, meta, True)


def array_stats(x, rows=64):
    """Fingerprint of a [n, d] transition array: sum / abs-sum / L2 over everything + `rows` strided rows."""
    x = np.asarray(x)
    x2 = x.reshape(len(x), -1).astype(np.float64)
    stride = max(1, len(x2) // rows)
    return np.concatenate([[x2.sum(), np.abs(x2).sum(), np.sqrt((x2 * x2).sum())], x2[::stride][:rows].ravel()])


def gen_rollout_cfg5(name, term, S=50_000, horizon=5, O=17, A=6, hidden=(256, 256), dyn_hidden=(200, 200, 200, 200), E=7,
                     n_elites=5, torch_seed=91, np_seed=92):
    """MOPOPolicy.rollout at BASELINE.json configs[4] size (50 000 start states x horizon 5, 7 members of 200 x 4, SAC actor
    256 x 2): too big to store, so the fixture holds the survivor counts per step (exact), per-array fingerprints, the
    seeds, and the small inputs are rebuilt from recipes.  The noise is re-drawn at test time from the same seeds in the
    reference's consumption order."""
    hidden, dyn_hidden = list(hidden), list(dyn_hidden)
    wds = [2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4][:len(dyn_hidden)] + [1e-4]
    model = EnsembleDynamicsModel(O, A, dyn_hidden, num_ensemble=E, num_elites=n_elites, weight_decays=wds, device="cpu")
    overwrite_params(model, 180)
    with torch.no_grad():
        model.max_logvar.fill_(0.5)
        model.min_logvar.fill_(-10.0)
        model.output_layer.weight.mul_(0.1)
        model.output_layer.bias[..., O + 1:] = -6.0
        model.set_elites([5, 0, 3, 6, 1][:n_elites])       # a non-trivial elite set
    tfn = {"halfcheetah": termination_fn_halfcheetah, "walker2d": termination_fn_walker2d}[term]
    data = make_dataset(S, O, A, seed=4)
    data["rewards"] = data["rewards"].reshape(-1, 1)
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), tfn, penalty_coef=0.5)
    inputs, _ = dyn.format_samples_for_training(data)
    dyn.scaler.fit(inputs)
    actor, c1, c2 = build_sac_like(O, A, hidden)
    for i, m in enumerate((actor, c1, c2)):
        overwrite_params(m, 181 + i)
    pol = MOPOPolicy(dyn, actor, c1, c2, torch.optim.Adam(actor.parameters(), lr=1e-4),
                     torch.optim.Adam(c1.parameters(), lr=3e-4), torch.optim.Adam(c2.parameters(), lr=3e-4), alpha=0.2)
    init = data["observations"].copy()
    if term == "walker2d":
        init[:, 0] = 1.4 + 0.35 * init[:, 0]      # heights around the walker2d band (0.8, 2.0): part of them terminates
        init[:, 1] = 0.3 * init[:, 1]
    torch.manual_seed(torch_seed)
    np.random.seed(np_seed)
    out, info = pol.rollout(init, horizon)
    counts, n_done, S_t = [], 0, S
    while n_done < len(out["obss"]):
        counts.append(S_t)
        term_t = out["terminals"][n_done:n_done + S_t]
        n_done += S_t
        S_t = int((~term_t).sum())
    assert n_done == info["num_transitions"]
    if term == "walker2d":      # distance of every decision to the termination thresholds (a flip would shift every later row)
        h, a = out["next_obss"][:, 0].astype(np.float64), out["next_obss"][:, 1].astype(np.float64)
        margin = min(np.abs(h - 0.8).min(), np.abs(h - 2.0).min(), np.abs(a - 1.0).min(), np.abs(a + 1.0).min())
        print(f"   seeds ({torch_seed}, {np_seed}): smallest distance to a termination threshold: {margin:.3e}", flush=True)
        if margin <= 1e-5:      # a state sits on a threshold: fp32-grade arithmetic could flip it; try the next seeds
            return gen_rollout_cfg5(name, term, S, horizon, O, A, hidden, dyn_hidden, E, n_elites, torch_seed + 2, np_seed + 2)
    store = {"counts": np.asarray(counts), "reward_mean": np.float64(info["reward_mean"]),
             "terminal_count": np.int64(out["terminals"].sum())}
    for k, v in out.items():
        store["outstats|" + k] = array_stats(v.astype(np.float64) if v.dtype == bool else v)
    meta = dict(algo="rollout_cfg5", O=O, A=A, hidden=hidden, dyn_hidden=dyn_hidden, E=E, n_elites=n_elites, S=S,
                horizon=horizon, term=term, penalty_coef=0.5, weight_decays=wds, data_seed=4, dyn_seed=180,
                elites=[5, 0, 3, 6, 1][:n_elites], param_seeds={"actor": 181, "critic1": 182, "critic2": 183},
                torch_seed=torch_seed, np_seed=np_seed, num_transitions=int(info["num_transitions"]))
    save(name, store, meta, False)
    print("   counts per step:", counts, " terminals:", int(out["terminals"].sum()))


def gen_cql_curve(name, n_steps=500, window=50, O=5, A=3, hidden=(32, 32, 32), B=64, N=4, n_data=4096):
    """Loss-CURVE fixture (north_star: "loss-curve parity with the reference"): the reference's CQLPolicy trained for
    `n_steps` on its own sampler and its own torch noise; stored are the per-window mean / std of every loss key for one
    run and, as the yardstick of what "statistically the same curve" means, the window means of a SECOND reference run
    that sees the same batches (same np.random seed) but other torch noise.  The engine draws Philox noise, so its curve
    is compared with run A the way run B is."""
    hidden = list(hidden)
    data = make_dataset(n_data, O, A, seed=0)
    hyper = dict(actor_lr=1e-4, critic_lr=3e-4, tau=0.005, gamma=0.99, cql_weight=5.0, temperature=1.0, max_q_backup=False,
                 deterministic_backup=True, with_lagrange=False, lagrange_threshold=10.0, cql_alpha_lr=3e-4, num_repeat_actions=N)

    def run(torch_seed):
        torch.manual_seed(0)
        actor, c1, c2 = build_sac_like(O, A, hidden)
        for i, m in enumerate((actor, c1, c2)):
            overwrite_params(m, 100 + i)
        log_alpha = torch.zeros(1, requires_grad=True)
        pol = CQLPolicy(actor, c1, c2, torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                        torch.optim.Adam(c1.parameters(), lr=hyper["critic_lr"]), torch.optim.Adam(c2.parameters(), lr=hyper["critic_lr"]),
                        action_space=gym.spaces.Box(-1, 1, (A,)), tau=hyper["tau"], gamma=hyper["gamma"],
                        alpha=(-A, log_alpha, torch.optim.Adam([log_alpha], lr=1e-4)), cql_weight=hyper["cql_weight"],
                        temperature=hyper["temperature"], max_q_backup=False, deterministic_backup=True, with_lagrange=False,
                        lagrange_threshold=10.0, cql_alpha_lr=3e-4, num_repeart_actions=N)
        pol.train()
        buf = ReplayBuffer(n_data, (O,), np.float32, A, np.float32, device="cpu")
        buf.load_dataset(data)
        np.random.seed(0)
        torch.manual_seed(torch_seed)
        rows = [pol.learn(buf.sample(B)) for _ in range(n_steps)]
        keys = sorted(rows[0])
        arr = np.asarray([[r[k] for k in keys] for r in rows], np.float64).reshape(n_steps // window, window, len(keys))
        return keys, arr.mean(1), arr.std(1)

    keys, mean_a, std_a = run(1)
    _, mean_b, _ = run(2)
    tol = 6.0 * std_a / np.sqrt(window) + 2e-3 * np.abs(mean_a) + 1e-4
    assert (np.abs(mean_b - mean_a) <= tol).all(), np.abs(mean_b - mean_a) / tol
    print("   second reference run vs first, worst |diff| / tol per key:",
          dict(zip(keys, np.round((np.abs(mean_b - mean_a) / tol).max(0), 3))))
    store = {"mean": mean_a, "std": std_a, "mean_other_noise": mean_b}
    meta = dict(algo="cql", O=O, A=A, hidden=hidden, B=B, N=N, n_steps=n_steps, window=window, n_data=n_data, data_seed=0,
                param_seeds={"actor": 100, "critic1": 101, "critic2": 102}, alpha_lr=1e-4, target_entropy=-A, hyper=hyper,
                np_seed=0, keys=keys)
    save(name, store, meta, False)


if __name__ == "__main__":
    only = sys.argv[1:]

    def run(fn, name, **kw):
        if not only or any(name.startswith(o) for o in only):
            fn(name, **kw)

    small = dict(O=5, A=3, hidden=[32, 32, 32])
    run(gen_cql, "cql_small", B=16, N=4, n_steps=3, with_lagrange=False, full_state=True, **small)
    run(gen_cql, "cql_small_lagrange", B=16, N=4, n_steps=3, with_lagrange=True, full_state=True, det_backup=False, **small)
    run(gen_cql, "cql_hc", O=17, A=6, hidden=[256, 256, 256], B=256, N=10, n_steps=2, with_lagrange=False, full_state=False)
    run(gen_cql, "cql_hc_lagrange", O=17, A=6, hidden=[256, 256, 256], B=256, N=10, n_steps=2, with_lagrange=True,
            full_state=False)
    run(gen_cql, "cql_small_maxq", B=16, N=4, n_steps=3, with_lagrange=True, full_state=True, max_q_backup=True, **small)
    run(gen_cql, "cql_hc_maxq", O=17, A=6, hidden=[256, 256, 256], B=256, N=10, n_steps=2, with_lagrange=False,
        full_state=False, max_q_backup=True)
    run(gen_cql, "cql_hopper", O=11, A=3, hidden=[256, 256, 256], B=256, N=10, n_steps=2, with_lagrange=False,
            full_state=False)
    run(gen_cql, "cql_hc_stochastic_backup", O=17, A=6, hidden=[256, 256, 256], B=256, N=10, n_steps=2, with_lagrange=False,
        full_state=False, det_backup=False)
    run(gen_combo, "combo_small_mix", n_real=10, n_fake=6, N=4, n_steps=3, rho_s="mix", with_lagrange=False, full_state=True,
        **small)
    run(gen_combo, "combo_small_model", n_real=9, n_fake=15, N=4, n_steps=3, rho_s="model", with_lagrange=True,
        full_state=True, det_backup=False, **small)
    run(gen_combo, "combo_hc", O=17, A=6, hidden=[256, 256, 256], n_real=128, n_fake=128, N=10, n_steps=2, rho_s="mix",
        with_lagrange=False, full_state=False)
    run(gen_combo, "combo_hc_model", O=17, A=6, hidden=[256, 256, 256], n_real=128, n_fake=128, N=10, n_steps=2,
        rho_s="model", with_lagrange=True, full_state=False)
    run(gen_sac, "sac_small", O=5, A=3, hidden=[32, 32], B=16, n_steps=3, full_state=True)
    run(gen_sac, "sac_hc", O=17, A=6, hidden=[256, 256], B=256, n_steps=2, full_state=False)
    run(gen_edac, "edac_small", O=5, A=3, hidden=[32, 32, 32], E=4, B=16, n_steps=3, full_state=True)
    run(gen_edac, "edac_small_maxq", O=5, A=3, hidden=[32, 32, 32], E=4, B=16, n_steps=3, full_state=True, max_q_backup=True)
    run(gen_edac, "edac_hc", O=17, A=6, hidden=[256, 256, 256], E=10, B=256, n_steps=2, full_state=False)
    run(gen_edac, "edac_hopper_e50", O=11, A=3, hidden=[256, 256, 256], E=50, B=256, n_steps=2, full_state=False, eta=1.0)
    run(gen_iql, "iql_small", O=5, A=3, hidden=[32, 32], B=16, n_steps=3, full_state=True)
    run(gen_iql, "iql_walker", O=17, A=6, hidden=[256, 256], B=256, n_steps=2, full_state=False)
    run(gen_iql, "iql_walker_b1024", O=17, A=6, hidden=[256, 256], B=1024, n_steps=2, full_state=False)
    run(gen_td3bc, "td3bc_small", O=5, A=3, hidden=[32, 32], B=16, n_steps=4, full_state=True)
    run(gen_td3bc, "td3bc_walker", O=17, A=6, hidden=[256, 256], B=256, n_steps=2, full_state=False)
    run(gen_td3bc, "td3bc_walker_b1024", O=17, A=6, hidden=[256, 256], B=1024, n_steps=2, full_state=False)
    run(gen_dynamics, "dynamics_small", O=5, A=3, hidden=[24, 24, 24, 24], E=3, n_elites=2, B=16, n_batches=3, S=64,
                 full_state=True, term="hopper")
    run(gen_dynamics, "dynamics_hc", O=17, A=6, hidden=[200, 200, 200, 200], E=7, n_elites=5, B=256, n_batches=2, S=512,
                 full_state=False, term="halfcheetah")
    run(gen_rollout, "rollout_small", O=5, A=3, hidden=[32, 32], dyn_hidden=[24, 24, 24, 24], E=3, n_elites=2, S=48,
                horizon=4, term="hopper")
    run(gen_rollout, "combo_rollout_uniform", O=5, A=3, hidden=[32, 32], dyn_hidden=[24, 24, 24, 24], E=3, n_elites=2, S=48,
                horizon=4, term="hopper", uniform=True)
    run(gen_dynamics_train, "dynamics_train_small", O=5, A=3, hidden=[24, 24, 24, 24], E=3, n_elites=2, n_data=640, B=32,
        max_epochs=6)
    run(gen_rollout_cfg5, "rollout_cfg5_hc", term="halfcheetah")
    run(gen_rollout_cfg5, "rollout_cfg5_walker", term="walker2d")
    run(gen_cql_curve, "cql_curve_small")
    run(gen_sample_next, "dynamics_sample_next_small", O=5, A=3, hidden=[24, 24, 24, 24], E=3, n_elites=2, S=32, num_samples=4,
        elites=[2, 0])
    run(gen_sample_next, "dynamics_sample_next_hc", O=17, A=6, hidden=[200, 200, 200, 200], E=7, n_elites=5, S=96,
        num_samples=10, elites=[6, 1, 3, 0, 4])
