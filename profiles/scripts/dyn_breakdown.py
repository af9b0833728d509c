import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from offlinerlkit_b200.modules import EnsembleDynamicsModel
from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.utils.scaler import StandardScaler
from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
O, A, B, E = 17, 6, 256, 7
torch.manual_seed(0)
model = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=E, num_elites=5, weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device="cuda:0")
dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
n = 256 * 20
x = torch.randn(n, O + A, device="cuda:0"); y = torch.randn(n, O + 1, device="cuda:0") * 0.1
idx = torch.randint(0, n, (E, n), device="cuda:0")
eng = dyn.engine
eng.learn(x, y, idx, B, 0.01)
plan = eng._learn_plans[B][0]
class Eng: pass
e = Eng(); e.plans = {"step": plan}; e.rt = eng.rt
br = bench.per_launch_breakdown(e)
print(f"{len(br)} launches, sum {sum(u for _, u in br):.1f} us")
for i, (l, u) in enumerate(br):
    print(f"{l:28s}{u:6.1f}", end="  ")
    if i % 4 == 3: print()
print()
