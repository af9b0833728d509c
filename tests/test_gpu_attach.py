"""GPU: attach mode (mode i) -- the engine behind objects built by the UNMODIFIED reference
(installed under baseline/_ref; skipped when it is absent), and the reference's own MFPolicyTrainer + Logger driving it."""
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import Golden, initial_state, assert_stats_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _ref():
    from baseline import reference_runner as rr
    ok, why = rr.available()
    if not ok:
        pytest.skip(f"reference not installed: {why}")
    import offlinerlkit
    return offlinerlkit


def _build_reference_policy(m):
    """The golden fixture's policy built from the REFERENCE's classes (as tests/golden/make_golden.py does on the CPU)."""
    _ref()
    import gym
    from offlinerlkit.nets import MLP
    from offlinerlkit.modules import ActorProb, Actor, Critic, EnsembleCritic, TanhDiagGaussian, DiagGaussian
    from offlinerlkit.policy import CQLPolicy, EDACPolicy, IQLPolicy, TD3BCPolicy, SACPolicy
    from offlinerlkit.utils.noise import GaussianNoise
    algo, O, A, hid, hy = m["algo"], m["O"], m["A"], m["hidden"], m.get("hyper", {})
    adam = lambda mod, lr: torch.optim.Adam(mod.parameters(), lr=lr)

    def alpha():
        la = torch.zeros(1, requires_grad=True, device=DEV)
        return (m["target_entropy"], la, torch.optim.Adam([la], lr=m["alpha_lr"]))

    def tanh_actor():
        bb = MLP(input_dim=O, hidden_dims=hid)
        return ActorProb(bb, TanhDiagGaussian(latent_dim=bb.output_dim, output_dim=A, unbounded=True, conditioned_sigma=True), DEV)

    if algo in ("cql", "sac"):
        actor, c1, c2 = tanh_actor(), Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
        opt = (adam(actor, hy["actor_lr"]), adam(c1, hy["critic_lr"]), adam(c2, hy["critic_lr"]))
        if algo == "sac":
            return SACPolicy(actor, c1, c2, *opt, tau=hy["tau"], gamma=hy["gamma"], alpha=alpha())
        return CQLPolicy(actor, c1, c2, *opt, action_space=gym.spaces.Box(-1, 1, (A,)), tau=hy["tau"], gamma=hy["gamma"],
                         alpha=alpha(), cql_weight=hy["cql_weight"], temperature=hy["temperature"], max_q_backup=hy["max_q_backup"],
                         deterministic_backup=hy["deterministic_backup"], with_lagrange=hy["with_lagrange"],
                         lagrange_threshold=hy["lagrange_threshold"], cql_alpha_lr=hy["cql_alpha_lr"],
                         num_repeart_actions=hy["num_repeat_actions"])
    if algo == "edac":
        actor = tanh_actor()
        critics = EnsembleCritic(O, A, hid, num_ensemble=m["E"], device=DEV)
        return EDACPolicy(actor, critics, adam(actor, hy["actor_lr"]), adam(critics, hy["critic_lr"]), tau=hy["tau"], gamma=hy["gamma"],
                          alpha=alpha(), max_q_backup=hy.get("max_q_backup", False), deterministic_backup=hy["deterministic_backup"],
                          eta=hy["eta"])
    if algo == "iql":
        bb = MLP(input_dim=O, hidden_dims=hid, dropout_rate=None)
        actor = ActorProb(bb, DiagGaussian(latent_dim=bb.output_dim, output_dim=A, unbounded=False, conditioned_sigma=False), DEV)
        q1, q2, v = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV), Critic(MLP(O, hid), DEV)
        return IQLPolicy(actor, q1, q2, v, adam(actor, hy["actor_lr"]), adam(q1, hy["critic_q_lr"]), adam(q2, hy["critic_q_lr"]),
                         adam(v, hy["critic_v_lr"]), action_space=gym.spaces.Box(-1, 1, (A,)), tau=hy["tau"], gamma=hy["gamma"],
                         expectile=hy["expectile"], temperature=hy["temperature"])
    if algo == "td3bc":
        actor = Actor(MLP(O, hid), A, device=DEV)
        c1, c2 = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
        return TD3BCPolicy(actor, c1, c2, adam(actor, hy["actor_lr"]), adam(c1, hy["critic_lr"]), adam(c2, hy["critic_lr"]),
                           tau=hy["tau"], gamma=hy["gamma"], max_action=hy["max_action"], exploration_noise=GaussianNoise(sigma=0.1),
                           policy_noise=hy["policy_noise"], noise_clip=hy["noise_clip"], update_actor_freq=hy["update_actor_freq"],
                           alpha=hy["alpha"], scaler=None)
    raise KeyError(algo)


@pytest.mark.parametrize("name", ["cql_small", "cql_hc", "sac_small", "edac_small", "iql_small", "td3bc_small"])
def test_attached_reference_objects_match_the_reference(name):
    """Reference policy + reference ReplayBuffer, attached: index stream and gather bit-exact, losses and parameters
    within 1e-4 of the reference's own CPU run (the golden fixture), parameters read through the REFERENCE's state_dict."""
    import offlinerlkit_b200 as orlk
    ref = _ref()
    from offlinerlkit.buffer import ReplayBuffer
    g = Golden(name)
    m = g.meta
    policy = _build_reference_policy(m)
    missing, unexpected = policy.load_state_dict({k: v.to(DEV) for k, v in initial_state(m).items()}, strict=False)
    assert not unexpected and all("saved_" in k for k in missing)
    policy.train()
    data = g.dataset()
    buf = ReplayBuffer(m["n_data"], (m["O"],), np.float32, m["A"], np.float32, device=DEV)
    buf.load_dataset(data)
    cls_p, cls_b = type(policy), type(buf)
    orlk.attach(policy, buf)
    assert isinstance(policy, cls_p) and isinstance(buf, cls_b) and type(policy).__name__ == "Attached" + cls_p.__name__
    np.random.seed(m["np_seed"])
    lr_atol = 2.5 * max(v for k, v in m["hyper"].items() if k.endswith("_lr"))
    for t in range(m["n_steps"]):
        batch = buf.sample(m["B"])
        assert np.array_equal(batch.indices.cpu().numpy(), g["idx"][t])
        for k, v in g.batch(t, data).items():
            assert torch.equal(batch[k].cpu().reshape(v.shape), v), k
        noise = g.noise(t) if any(k.startswith(f"noise{t}|") for k in g.z.files) else None
        out = policy.learn(batch, noise=noise) if noise is not None else policy.learn(batch)
        ref_l = g.losses(t)
        assert out.keys() == ref_l.keys()
        for k in ref_l:
            assert abs(out[k] - ref_l[k]) <= TOL * max(1.0, abs(ref_l[k])), (t, k, out[k], ref_l[k])
        sd = {k: v.detach().cpu() for k, v in policy.state_dict().items()}
        assert_stats_close(sd, g.group(f"stats{t}"), tol=TOL, lr_atol=lr_atol * (1 + 0.8 * t))
    if getattr(policy, "_is_auto_alpha", False):
        assert torch.is_tensor(policy._alpha) and float(policy._alpha) == pytest.approx(out["alpha"])


class _StubEnv:
    """What MFPolicyTrainer._evaluate needs of a gym env (mf_policy_trainer.py:92-125)."""

    def __init__(self, O, A):
        self.O, self.A, self.t = O, A, 0

    def reset(self):
        self.t = 0
        return np.zeros(self.O, np.float32)

    def step(self, action):
        self.t += 1
        return np.full(self.O, 0.1 * self.t, np.float32), 1.0, self.t >= 5, {}

    def get_normalized_score(self, r):
        return r / 100.0


def test_unmodified_reference_trainer_drives_the_engine(tmp_path):
    """The reference's MFPolicyTrainer and Logger, unmodified, for 2 epochs x 25 steps over the attached policy + buffer;
    the logged loss means equal a plain sample/learn loop of the same 50 steps on a second attached policy."""
    import offlinerlkit_b200 as orlk
    _ref()
    from offlinerlkit.buffer import ReplayBuffer
    from offlinerlkit.policy_trainer import MFPolicyTrainer
    from offlinerlkit.utils.logger import Logger
    g = Golden("cql_small")
    m = g.meta
    data = g.dataset()

    def make():
        torch.manual_seed(0)
        pol = _build_reference_policy(m)
        pol.load_state_dict({k: v.to(DEV) for k, v in initial_state(m).items()}, strict=False)
        buf = ReplayBuffer(m["n_data"], (m["O"],), np.float32, m["A"], np.float32, device=DEV)
        buf.load_dataset(data)
        orlk.attach(pol, buf)
        pol.engine(m["B"]).seed = 123
        return pol, buf

    pol, buf = make()
    logger = Logger(str(tmp_path), {"consoleout_backup": "stdout", "policy_training_progress": "csv"})
    trainer = MFPolicyTrainer(policy=pol, eval_env=_StubEnv(m["O"], m["A"]), buffer=buf, logger=logger, epoch=2, step_per_epoch=25,
                              batch_size=m["B"], eval_episodes=2)
    np.random.seed(5)
    trainer.train()
    import csv
    rows = list(csv.DictReader(open(os.path.join(str(tmp_path), "record", "policy_training_progress.csv"))))
    assert len(rows) == 2 and "loss/critic1" in rows[0] and "eval/normalized_episode_reward" in rows[0]
    assert os.path.exists(os.path.join(str(tmp_path), "checkpoint", "policy.pth"))

    pol2, buf2 = make()
    np.random.seed(5)
    pol2.train()
    means = []
    for e in range(2):
        acc = {}
        for _ in range(25):
            for k, v in pol2.learn(buf2.sample(m["B"])).items():
                acc.setdefault(k, []).append(v)
        means.append({k: float(np.mean(v)) for k, v in acc.items()})
    for e in range(2):
        for k, v in means[e].items():
            assert float(rows[e][k]) == pytest.approx(v, rel=1e-6, abs=1e-7), (e, k)
