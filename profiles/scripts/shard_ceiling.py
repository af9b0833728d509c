"""Upper bound of member sharding (EDAC critics and MOPO dynamics training): the time of one
step / one mini-batch on ONE GPU as a function of the number of ensemble members it holds.  A rank of a G-way
member-sharded job runs E/G members plus the per-step collectives; the figures here are its compute alone, so
t(E) / t(E/G) is the speed-up sharding could reach before any communication.  Usage: shard_ceiling.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from tests.gpu_common import build_policy, make_buffer
from tests.helpers import Golden

out = {"edac_step_us": {}, "dynamics_batch_us": {}}
g = Golden("edac_hc")
for E in (50, 10, 5, 3, 2):        # 50 = the reference's hopper setting; the diversity term needs E >= 2
    m = dict(g.meta, E=E)
    torch.manual_seed(0)
    np.random.seed(0)
    pol = build_policy(m)
    pol.train()
    buf, _ = make_buffer(g)
    for _ in range(20):
        pol.learn(buf.sample(m["B"]))
    eng = pol._engine
    plan = eng.plans[sorted(eng.plans)[0]]
    s = torch.cuda.ExternalStream(eng.rt.cur.value) if eng.rt.cur.value else torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 500
    with torch.cuda.stream(s):
        plan.launch()
        ev0.record(s)
        for _ in range(N):
            plan.launch()
        ev1.record(s)
    torch.cuda.synchronize()
    out["edac_step_us"][E] = round(ev0.elapsed_time(ev1) * 1e3 / N, 1)
    print("edac E", E, out["edac_step_us"][E], flush=True)

from offlinerlkit_b200.modules import EnsembleDynamicsModel
from offlinerlkit_b200.dynamics import EnsembleDynamics
from offlinerlkit_b200.utils.scaler import StandardScaler
from offlinerlkit_b200.utils.termination_fns import termination_fn_halfcheetah
O, A, B = 17, 6, 256
for E in (7, 4, 2, 1):
    torch.manual_seed(0)
    model = EnsembleDynamicsModel(O, A, [200, 200, 200, 200], num_ensemble=E, num_elites=max(1, E - 2),
                                  weight_decays=[2.5e-5, 5e-5, 7.5e-5, 7.5e-5, 1e-4], device="cuda:0")
    dyn = EnsembleDynamics(model, torch.optim.Adam(model.parameters(), lr=1e-3), StandardScaler(), termination_fn_halfcheetah)
    n = 256 * 200
    x = torch.randn(n, O + A, device="cuda:0")
    y = torch.randn(n, O + 1, device="cuda:0") * 0.1
    idx = torch.randint(0, n, (E, n), device="cuda:0")
    eng = dyn.engine
    eng.learn(x, y, idx, B, 0.01)
    torch.cuda.synchronize()
    s = torch.cuda.ExternalStream(eng.rt.cur.value) if eng.rt.cur.value else torch.cuda.current_stream()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        ev0.record(s)
    eng.learn(x, y, idx, B, 0.01)
    with torch.cuda.stream(s):
        ev1.record(s)
    torch.cuda.synchronize()
    out["dynamics_batch_us"][E] = round(ev0.elapsed_time(ev1) * 1e3 / (n // B), 1)
    print("dynamics E", E, out["dynamics_batch_us"][E], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
