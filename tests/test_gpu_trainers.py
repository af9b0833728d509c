"""GPU: the trainer loops drive the engine through the reference's contracts (buffer.sample -> policy.learn -> loss dict,
periodic rollouts into the fake buffer, evaluation via select_action, state_dict checkpoints)."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import Golden, initial_state

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class FakeEnv:
    def __init__(self, obs_dim, act_dim, horizon=5):
        self.o, self.a, self.h, self.t = obs_dim, act_dim, horizon, 0

    def reset(self):
        self.t = 0
        return np.zeros(self.o, np.float32)

    def step(self, action):
        assert action.shape == (self.a,)
        self.t += 1
        return np.full(self.o, 0.1 * self.t, np.float32), 1.0, self.t >= self.h, {}

    def get_normalized_score(self, x):
        return x / 10.0


def test_mf_trainer_runs_cql(tmp_path):
    from tests.gpu_common import build_policy, load_state, make_buffer
    from offlinerlkit_b200.policy_trainer import MFPolicyTrainer
    from offlinerlkit_b200.utils.logger import Logger
    g = Golden("cql_small")
    m = g.meta
    policy = build_policy(m, DEV)
    load_state(policy, initial_state(m))
    buf, _ = make_buffer(g, DEV)
    logger = Logger(str(tmp_path))
    logger.quiet = True
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(policy.actor_optim, 2)
    trainer = MFPolicyTrainer(policy, FakeEnv(m["O"], m["A"]), buf, logger, epoch=2, step_per_epoch=20, batch_size=m["B"],
                              eval_episodes=2, lr_scheduler=sched)
    before = {k: v.detach().clone() for k, v in policy.state_dict().items()}
    out = trainer.train()
    assert np.isfinite(out["last_10_performance"])
    ck = torch.load(os.path.join(logger.model_dir, "policy.pth"))
    assert set(ck) == set(before)
    assert any(not torch.equal(ck[k].cpu(), before[k].cpu()) for k in ck)
    assert policy._engine.group_steps()[0] == 40               # 40 actor Adam steps were applied on the device
    # the lr scheduler's value after the first epoch (cosine, T_max=2: 1e-4 -> 5e-5) reached the device-side Adam table
    assert policy._engine._group_lr[policy._engine.g_actor] == pytest.approx(0.5 * m["hyper"]["actor_lr"])


@pytest.mark.parametrize("algo", ["mopo", "combo", "combo_uniform_model"])
def test_mb_trainer_runs_mopo(tmp_path, algo):
    """run_example/run_mopo.py and run_combo.py in miniature: dynamics.train -> MBPolicyTrainer (rollouts into the fake
    buffer, policy.learn on {"real", "fake"} batches)."""
    from offlinerlkit_b200.nets import MLP
    from offlinerlkit_b200.modules import ActorProb, Critic, TanhDiagGaussian, EnsembleDynamicsModel
    from offlinerlkit_b200.policy import MOPOPolicy, COMBOPolicy
    from offlinerlkit_b200.dynamics import EnsembleDynamics
    from tests.gpu_common import Box
    from offlinerlkit_b200.buffer import ReplayBuffer
    from offlinerlkit_b200.policy_trainer import MBPolicyTrainer
    from offlinerlkit_b200.utils.logger import Logger
    from offlinerlkit_b200.utils.scaler import StandardScaler
    from offlinerlkit_b200.utils.termination_fns import get_termination_fn
    from offlinerlkit_b200.synthetic import make_dataset
    O, A, hid = 5, 3, [32, 32]
    torch.manual_seed(0)
    np.random.seed(0)
    data = make_dataset(2000, O, A, seed=3)
    bb = MLP(O, hid)
    actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), DEV)
    c1, c2 = Critic(MLP(O + A, hid), DEV), Critic(MLP(O + A, hid), DEV)
    adam = lambda mod, lr: torch.optim.Adam(mod.parameters(), lr=lr)
    model = EnsembleDynamicsModel(O, A, [24, 24], num_ensemble=3, num_elites=2, weight_decays=[2.5e-5, 5e-5, 7.5e-5], device=DEV)
    dyn = EnsembleDynamics(model, adam(model, 1e-3), StandardScaler(), get_termination_fn("halfcheetah-medium-v2"), penalty_coef=0.5)
    la = torch.zeros(1, requires_grad=True, device=DEV)
    opt = (adam(actor, 1e-4), adam(c1, 3e-4), adam(c2, 3e-4))
    if algo == "mopo":
        pol = MOPOPolicy(dyn, actor, c1, c2, *opt, alpha=(-A, la, torch.optim.Adam([la], lr=1e-4)))
    else:
        pol = COMBOPolicy(dyn, actor, c1, c2, *opt, action_space=Box(-1, 1, (A,)),
                          alpha=(-A, la, torch.optim.Adam([la], lr=1e-4)), cql_weight=5.0, with_lagrange=False,
                          num_repeart_actions=4, uniform_rollout=algo.endswith("uniform_model"),
                          rho_s="model" if algo.endswith("model") else "mix")
    real = ReplayBuffer(2000, (O,), np.float32, A, np.float32, device=DEV)
    real.load_dataset(data)
    fake = ReplayBuffer(64 * 2 * 5, (O,), np.float32, A, np.float32, device=DEV)
    logger = Logger(str(tmp_path))
    logger.quiet = True
    dyn.train(real.sample_all(), logger, max_epochs=2, max_epochs_since_update=5)
    assert os.path.exists(os.path.join(logger.model_dir, "dynamics.pth"))
    assert len(set(model.elites.tolist())) == 2
    trainer = MBPolicyTrainer(pol, FakeEnv(O, A), real, fake, logger, rollout_setting=(10, 64, 2), epoch=1, step_per_epoch=25,
                              batch_size=32, real_ratio=0.25, eval_episodes=1)
    trainer.train()
    assert fake._size > 0 and np.isfinite(fake.rewards[:fake._size]).all()
    assert os.path.exists(os.path.join(logger.model_dir, "policy.pth"))
