"""Runs the UNMODIFIED reference (zhaoyizhou1123/OfflineRL-Kit) for bench.py's reference arm and denominators.

The reference is installed once, in the authoring container, into the git-ignored ``baseline/_ref`` with

    cp -r /root/reference /tmp/ref_copy          # the reference tree is read-only and pip builds in-tree
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
        --target baseline/_ref /tmp/ref_copy
    cp -r --update=none /tmp/ref_copy/offlinerlkit baseline/_ref/     # setup.py's find_packages() skips policy/others (no
                                                                      # __init__.py), which policy/__init__.py imports

(``--no-deps``: gym / ray / d4rl are not in the image; the five import-time stubs under ``tests/golden/_stubs`` --
gym, gymnasium, diffusers, wandb, matplotlib -- stand in for packages the hot path never calls)
``baseline/_ref`` travels to the GPU box with the gpurun snapshot.  Nothing here touches the repo's own engine: the
policy, the buffer and the optimisers are the reference's classes, built exactly as ``run_example/run_cql.py:72-139``
builds them, and a step is the stock ``buffer.sample(256)`` + ``policy.learn(batch)``.
"""
import os
import random
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(ROOT, "tests", "golden", "_stubs")


def available():
    """(True, "") when the installed reference can be imported, else (False, reason)."""
    if not os.path.isdir(os.path.join(REF, "offlinerlkit")):
        return False, "baseline/_ref/offlinerlkit is absent (install recipe in baseline/reference_runner.py)"
    try:
        _import()
    except Exception as ex:          # pragma: no cover - depends on the box
        return False, f"reference import failed: {ex!r}"
    return True, ""


def _import():
    for p in (REF, STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import offlinerlkit            # noqa: F401
    assert os.path.realpath(os.path.dirname(offlinerlkit.__file__)).startswith(os.path.realpath(REF)), \
        "offlinerlkit must come from baseline/_ref"
    from offlinerlkit.nets import MLP
    from offlinerlkit.modules import ActorProb, Critic, TanhDiagGaussian
    from offlinerlkit.buffer import ReplayBuffer
    from offlinerlkit.policy import CQLPolicy
    import gym
    return MLP, ActorProb, Critic, TanhDiagGaussian, ReplayBuffer, CQLPolicy, gym


def build_cql(device: str, seed: int, n_data: int, obs_dim: int, act_dim: int, hidden, hyper: dict, alpha_lr: float):
    """run_example/run_cql.py:72-139 on the synthetic dataset of BASELINE.md section 3."""
    MLP, ActorProb, Critic, TanhDiagGaussian, ReplayBuffer, CQLPolicy, gym = _import()
    sys.path.insert(0, ROOT)
    from offlinerlkit_b200.synthetic import make_dataset
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    actor_backbone = MLP(input_dim=obs_dim, hidden_dims=hidden)
    c1b = MLP(input_dim=obs_dim + act_dim, hidden_dims=hidden)
    c2b = MLP(input_dim=obs_dim + act_dim, hidden_dims=hidden)
    dist = TanhDiagGaussian(latent_dim=getattr(actor_backbone, "output_dim"), output_dim=act_dim, unbounded=True,
                            conditioned_sigma=True)
    actor, critic1, critic2 = ActorProb(actor_backbone, dist, device), Critic(c1b, device), Critic(c2b, device)
    log_alpha = torch.zeros(1, requires_grad=True, device=device)
    alpha = (-act_dim, log_alpha, torch.optim.Adam([log_alpha], lr=alpha_lr))
    policy = CQLPolicy(actor, critic1, critic2,
                       torch.optim.Adam(actor.parameters(), lr=hyper["actor_lr"]),
                       torch.optim.Adam(critic1.parameters(), lr=hyper["critic_lr"]),
                       torch.optim.Adam(critic2.parameters(), lr=hyper["critic_lr"]),
                       action_space=gym.spaces.Box(-1, 1, (act_dim,)), tau=hyper["tau"], gamma=hyper["gamma"], alpha=alpha,
                       cql_weight=hyper["cql_weight"], temperature=hyper["temperature"],
                       max_q_backup=hyper["max_q_backup"], deterministic_backup=hyper["deterministic_backup"],
                       with_lagrange=hyper["with_lagrange"], lagrange_threshold=hyper["lagrange_threshold"],
                       cql_alpha_lr=hyper["cql_alpha_lr"], num_repeart_actions=hyper["num_repeat_actions"])
    policy.train()
    buf = ReplayBuffer(buffer_size=n_data, obs_shape=(obs_dim,), obs_dtype=np.float32, action_dim=act_dim,
                       action_dtype=np.float32, device=device)
    buf.load_dataset(make_dataset(n_data, obs_dim, act_dim, seed=0))
    return policy, buf


def time_cql(device: str, steps: int, warmup: int, budget_s: float, n_data: int, batch: int, **build_kw):
    """steps/s of the stock ``sample`` + ``learn`` loop (wall clock, device synchronised on both sides)."""
    policy, buf = build_cql(device, 0, n_data, **build_kw)
    cuda = device != "cpu"
    for _ in range(warmup):
        policy.learn(buf.sample(batch))
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    loss = None
    while done < steps:
        loss = policy.learn(buf.sample(batch))
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    if cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return done / dt, done, dt, loss
