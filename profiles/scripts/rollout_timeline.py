"""Kernel timeline of one MOPO rollout (config 5 shape) under torch.profiler: per-kernel totals and the device idle
time.  Profiling aid only.  Usage: python profiles/scripts/rollout_timeline.py"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.modules import ActorProb, Critic, EnsembleDynamicsModel, TanhDiagGaussian
from offlinerlkit_b200.nets import MLP
from offlinerlkit_b200.policy import MOPOPolicy
from offlinerlkit_b200.utils import termination_fns as T
from offlinerlkit_b200.utils.scaler import StandardScaler

dev = "cuda:0"
O, A, S, H = 17, 6, 50_000, 5
torch.manual_seed(0)
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
adam = lambda m: torch.optim.Adam(m.parameters(), lr=1e-4)
pol = MOPOPolicy(dyn, actor, c1, c2, adam(actor), adam(c1), adam(c2), alpha=0.2)
init = np.random.default_rng(1).standard_normal((S, O), dtype=np.float32)
for _ in range(3):
    pol.rollout(init, H)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pol.rollout(init, H)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
tot = defaultdict(lambda: [0, 0.0])
for e in ev:
    tot[e.name[:70]][0] += 1
    tot[e.name[:70]][1] += e.time_range.end - e.time_range.start
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[1] for v in tot.values())
print(f"{len(ev)} device activities; span {span / 1e3:.2f} ms; busy {busy / 1e3:.2f} ms; timing {pol._roll.last_timing}")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t / 1e3:8.3f} ms  x{n:4d}  {k}")
