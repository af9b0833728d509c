"""One MBPolicyTrainer cycle at MOPO's defaults (mb_policy_trainer.py:66-89, run_mopo.py: rollout_freq 1000,
rollout_batch_size 50 000, rollout_length 5, batch 256, real_ratio 0.05): a model rollout from 50 000 buffer states,
its hand-off to the model buffer, then 1000 SAC steps on {"real", "fake"} batches.  Wall-clock split per cycle.
Shapes of config 5 (O=17, A=6, E=7, hidden 200x4, actor/critics 256x2), synthetic data, random-init weights.
Usage: python profiles/scripts/mopo_loop.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from offlinerlkit_b200.buffer import ReplayBuffer
from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.modules import ActorProb, Critic, EnsembleDynamicsModel, TanhDiagGaussian
from offlinerlkit_b200.nets import MLP
from offlinerlkit_b200.policy import MOPOPolicy
from offlinerlkit_b200.synthetic import make_dataset
from offlinerlkit_b200.utils import termination_fns as T
from offlinerlkit_b200.utils.scaler import StandardScaler

dev = "cuda:0"
O, A, N = 17, 6, 200_000
FREQ, RB, RL, B, RATIO = 1000, 50_000, 5, 256, 0.05
torch.manual_seed(0)
np.random.seed(0)
model = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=7, num_elites=5,
                              weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device=dev)
with torch.no_grad():
    for k, v in model.state_dict().items():
        if "backbones.3" in k or "output" in k:
            v.mul_(0.1)
dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3),
                       StandardScaler(np.zeros((1, O + A), np.float32), np.ones((1, O + A), np.float32)),
                       T.termination_fn_halfcheetah, penalty_coef=0.5)
dyn.rng = "device"
bb = MLP(O, [256, 256])
actor = ActorProb(bb, TanhDiagGaussian(bb.output_dim, A, unbounded=True, conditioned_sigma=True), dev)
c1, c2 = Critic(MLP(O + A, [256, 256]), dev), Critic(MLP(O + A, [256, 256]), dev)
adam = lambda m, lr: torch.optim.Adam(m.parameters(), lr=lr)
la = torch.zeros(1, requires_grad=True, device=dev)
pol = MOPOPolicy(dyn, actor, c1, c2, adam(actor, 1e-4), adam(c1, 3e-4), adam(c2, 3e-4),
                 alpha=(-A, la, torch.optim.Adam([la], lr=1e-4)))
pol.train()
real = ReplayBuffer(N, (O,), np.float32, A, np.float32, device=dev)
real.load_dataset(make_dataset(N, O, A, seed=0))
fake = ReplayBuffer(RB * RL * 5, (O,), np.float32, A, np.float32, device=dev)      # model_retain_epochs = 5


def cycle(device_handoff: bool):
    t0 = time.perf_counter()
    init = real.sample(RB)["observations"].cpu().numpy()
    if device_handoff:
        tr, info = pol.rollout(init, RL, device_out=True)
    else:
        tr, info = pol.rollout(init, RL)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    fake.add_batch(**tr)
    torch.cuda.synchronize()
    t_add = time.perf_counter()
    real_n = int(B * RATIO)
    loss = pol.learn({"real": real.sample(real_n), "fake": fake.sample(B - real_n)})      # first step pays the mirror sync
    t2 = time.perf_counter()
    for _ in range(FREQ - 1):
        loss = pol.learn({"real": real.sample(real_n), "fake": fake.sample(B - real_n)})
    t3 = time.perf_counter()
    return {"rollout_ms": 1e3 * (t1 - t0), "add_batch_ms": 1e3 * (t_add - t1), "handoff_plus_first_step_ms": 1e3 * (t2 - t1), "learn_999_ms": 1e3 * (t3 - t2),
            "cycle_ms": 1e3 * (t3 - t0), "transitions": int(info["num_transitions"]), "finite": bool(np.isfinite(list(loss.values())).all())}


out = {}
for name, flag in (("host_handoff", False), ("device_handoff", True)):
    cycle(flag)
    cycle(flag)
    runs = [cycle(flag) for _ in range(3)]
    out[name] = {k: (round(float(np.median([r[k] for r in runs])), 2) if isinstance(runs[0][k], float) else runs[0][k])
                 for k in runs[0]}
    print(name, out[name], "add_batch per run:", [round(r["add_batch_ms"], 1) for r in runs], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
